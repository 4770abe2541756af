// stand-in: MeasurementBuffer (declared, never instantiated under oracle/_ref) holds a reference to one
#ifndef STVL_REF_SHIM_TF2_BUFFER_H_
#define STVL_REF_SHIM_TF2_BUFFER_H_
namespace tf2_ros { class Buffer {}; }
#endif
