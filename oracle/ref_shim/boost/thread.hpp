// stand-in: boost::mutex / unique_lock / recursive_mutex as used by the reference (one lock per public grid call)
#ifndef STVL_REF_SHIM_BOOST_THREAD_HPP_
#define STVL_REF_SHIM_BOOST_THREAD_HPP_
#include <mutex>
namespace boost
{
typedef std::mutex mutex;
typedef std::recursive_mutex recursive_mutex;
template <typename M> using unique_lock = std::unique_lock<M>;
}
#endif
