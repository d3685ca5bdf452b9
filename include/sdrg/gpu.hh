// sdrg/gpu.hh -- glue between the node classes and the C ABI (include/sdrg.h).
#ifndef SDRG_GPU_HH
#define SDRG_GPU_HH

#include "../sdrg.h"
#include "buffer.hh"
#include "exception.hh"

namespace sdr {
namespace gpu {

/** Maps an sdrg status to the reference's exception types (src/exception.hh:10-45). */
inline void check(int rc) {
  if (SDRG_OK == rc) return;
  if (SDRG_ERR_CONFIG == rc) { ConfigError err; err << sdrg_last_error(); throw err; }
  RuntimeError err; err << sdrg_last_error(); throw err;
}

inline void *stream() { void *s = 0; check(sdrg_stream_default(&s)); return s; }

/** Device address holding the bytes of `b` (upload on demand; foreign memory is staged through
 * thread-private scratch). */
inline const void *deviceInput(const RawBuffer &b, void *st) {
  void *d = 0;
  check(sdrg_buffer_to_device(b.data(), b.bytesLen(), st, &d));
  if (d) return d;
  check(sdrg_scratch(b.bytesLen(), &d));
  check(sdrg_memcpy_h2d_async(d, b.data(), b.bytesLen(), st));
  return d;
}

/** Device address a node may write its result for `b` to, or 0 if `b` has no device mirror. */
inline void *deviceOutput(const RawBuffer &b) {
  void *d = 0;
  check(sdrg_buffer_device_ptr(b.data(), &d));
  return d;
}

/** Declares that [b.data(), +bytes) was just produced on the device (stream order). */
inline void publish(const RawBuffer &b, size_t bytes, void *st) { check(sdrg_buffer_mark_device_valid(b.data(), bytes, st)); }

}  // namespace gpu
}  // namespace sdr
#endif
