#include "boost/thread.hpp"
