// Stand-in for sensor_msgs/point_cloud2_iterator.hpp: named-field strided iterators over PointCloud2::data (the reference
// reads x, y, z through PointCloud2ConstIterator<float>, .cpp:278-280) and the modifier it uses to emit its occupancy cloud
// (.cpp:352-359).  Same semantics as the ROS header: the iterator starts at the field's offset and advances by point_step;
// end() is data + width*height*point_step + offset.
#ifndef STVL_REF_SHIM_SENSOR_MSGS_POINT_CLOUD2_ITERATOR_HPP_
#define STVL_REF_SHIM_SENSOR_MSGS_POINT_CLOUD2_ITERATOR_HPP_
#include <cstdarg>
#include <cstring>
#include <stdexcept>
#include "sensor_msgs/msg/point_cloud2.hpp"
namespace sensor_msgs
{
template <typename T, typename CloudT, typename ByteT>
class PointCloud2IteratorBase
{
public:
  PointCloud2IteratorBase(CloudT & cloud, const std::string & field)
  {
    uint32_t off = 0;
    bool found = false;
    for (const auto & f : cloud.fields) if (f.name == field) { off = f.offset; found = true; }
    if (!found) throw std::runtime_error("Field " + field + " does not exist");
    step_ = cloud.point_step;
    ByteT * base = cloud.data.data();
    ptr_ = base + off;
    end_ = base + (size_t)cloud.width * cloud.height * cloud.point_step + off;
  }
  T & operator*() const { return *reinterpret_cast<T *>(const_cast<uint8_t *>(reinterpret_cast<const uint8_t *>(ptr_))); }
  PointCloud2IteratorBase & operator++() { ptr_ += step_; return *this; }
  bool operator!=(const PointCloud2IteratorBase & o) const { return ptr_ != o.ptr_; }
  PointCloud2IteratorBase end() const { PointCloud2IteratorBase r(*this); r.ptr_ = end_; return r; }
private:
  ByteT * ptr_ = nullptr, * end_ = nullptr;
  uint32_t step_ = 0;
};
template <typename T> using PointCloud2Iterator = PointCloud2IteratorBase<T, msg::PointCloud2, uint8_t>;
template <typename T> using PointCloud2ConstIterator = PointCloud2IteratorBase<const T, const msg::PointCloud2, const uint8_t>;

class PointCloud2Modifier
{
public:
  explicit PointCloud2Modifier(msg::PointCloud2 & c) : c_(c) {}
  // setPointCloud2Fields(n, name, count, datatype, ...): consecutive fields, then resize to width*height points
  void setPointCloud2Fields(int n_fields, ...)
  {
    c_.fields.clear();
    va_list ap;
    va_start(ap, n_fields);
    uint32_t off = 0;
    for (int i = 0; i < n_fields; ++i) {
      msg::PointField f;
      f.name = va_arg(ap, const char *);
      f.count = (uint32_t)va_arg(ap, int);
      f.datatype = (uint8_t)va_arg(ap, int);
      f.offset = off;
      off += f.count * 4;
      c_.fields.push_back(f);
    }
    va_end(ap);
    finish(off);
  }
  // setPointCloud2FieldsByString(1, "xyz"): x, y, z FLOAT32 plus 4 bytes of padding (point_step 16)
  void setPointCloud2FieldsByString(int, ...)
  {
    c_.fields.clear();
    const char * names[3] = {"x", "y", "z"};
    for (uint32_t i = 0; i < 3; ++i) { msg::PointField f; f.name = names[i]; f.offset = 4 * i; f.datatype = msg::PointField::FLOAT32; f.count = 1; c_.fields.push_back(f); }
    finish(16);
  }
private:
  void finish(uint32_t step)
  {
    c_.point_step = step;
    c_.row_step = step * c_.width;
    c_.data.assign((size_t)c_.row_step * c_.height, 0);
  }
  msg::PointCloud2 & c_;
};
}  // namespace sensor_msgs
#endif
