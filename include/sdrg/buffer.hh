// sdrg/buffer.hh -- the Buffer family of the host-side mirror.
//
// Interface and observable behaviour mirrored from src/buffer.hh:19-541 / src/buffer.cc (checked by
// tests/cpp/buffer_test.cc, a restatement of test/buffertest.cc:9-122):
//   * RawBuffer(size_t N, BufferOwner*) allocates storage with reference count 1; copies, views and
//     assignment do NOT count; ref()/unref() are explicit; at count 1 the owner's bufferUnused()
//     fires, at 0 the storage is released and the view becomes empty; isUnused() <=> count == 1 or
//     the view is unowned.
//   * Buffer<T>(const RawBuffer&) reinterprets the bytes; operator[] is a raw host dereference.
// What changed underneath: the storage of owned buffers is pinned host memory with a device mirror
// (sdrg_buffer_alloc) so that GPU nodes can consume and produce it without staging; the device-side
// validity of the bytes is tracked by the library, keyed by the host address.  Small buffers and
// hosts without a CUDA device use the heap.
#ifndef SDRG_BUFFER_HH
#define SDRG_BUFFER_HH

#include <atomic>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <map>
#include <ostream>
#include <vector>
#include <inttypes.h>

#include "../sdrg.h"
#include "exception.hh"

namespace sdr {

class RawBuffer;

/** Anyone handing out pooled buffers: told when only its own reference is left. */
class BufferOwner {
public:
  virtual ~BufferOwner() {}
  virtual void bufferUnused(const RawBuffer &buffer) = 0;
};

namespace detail {
/** Shared bookkeeping of one allocation. */
struct Storage {
  std::atomic<int> refs;
  BufferOwner *owner;
  bool managed;            // allocated by sdrg_buffer_alloc (pinned + device mirror)
  static constexpr size_t kManagedThreshold = 4096;   // smaller buffers never travel to the GPU in bulk
  static char *allocate(size_t bytes, bool &managed, bool force_device) {
    managed = false;
    if (bytes >= kManagedThreshold || (force_device && bytes > 0)) {
      void *p = 0;
      if (SDRG_OK == sdrg_buffer_alloc(bytes, &p) && p) { managed = true; return (char *)p; }
    }
    return (char *)std::malloc(bytes ? bytes : 1);
  }
  static void release(char *p, bool managed) {
    if (!p) return;
    if (managed) sdrg_buffer_free(p); else std::free(p);
  }
};
}  // namespace detail

class RawBuffer {
public:
  RawBuffer() : _ptr(0), _storage_size(0), _b_offset(0), _b_length(0), _st(0) {}
  /** Wraps memory the buffer does not own (never counted, never freed). */
  RawBuffer(char *data, size_t offset, size_t len)
    : _ptr(data), _storage_size(offset + len), _b_offset(offset), _b_length(len), _st(0) {}
  /** Allocates N bytes; the new buffer holds the one and only reference.  `device_backed` asks for
   * pinned + device-mirrored storage whatever the size (GPU nodes do so for their output buffers). */
  RawBuffer(size_t N, BufferOwner *owner = 0, bool device_backed = false)
    : _ptr(0), _storage_size(0), _b_offset(0), _b_length(0), _st(0) {
    bool managed = false;
    char *p = detail::Storage::allocate(N, managed, device_backed);
    if (!p) return;
    _st = new detail::Storage();
    _st->refs.store(1); _st->owner = owner; _st->managed = managed;
    _ptr = p; _storage_size = N; _b_length = N;
  }
  RawBuffer(const RawBuffer &o)
    : _ptr(o._ptr), _storage_size(o._storage_size), _b_offset(o._b_offset), _b_length(o._b_length), _st(o._st) {}
  /** A view of len bytes starting offset bytes into other's view. */
  RawBuffer(const RawBuffer &o, size_t offset, size_t len)
    : _ptr(o._ptr), _storage_size(o._storage_size), _b_offset(o._b_offset + offset), _b_length(len), _st(o._st) {}
  virtual ~RawBuffer() {}

  const RawBuffer &operator=(const RawBuffer &o) {
    _ptr = o._ptr; _storage_size = o._storage_size; _b_offset = o._b_offset; _b_length = o._b_length; _st = o._st;
    return *this;
  }

  inline char *ptr() const { return _ptr; }
  inline char *data() const { return _ptr + _b_offset; }
  inline size_t bytesOffset() const { return _b_offset; }
  inline size_t bytesLen() const { return _b_length; }
  inline size_t storageSize() const { return _storage_size; }
  inline bool isEmpty() const { return 0 == _ptr; }

  void ref() const { if (_st) _st->refs.fetch_add(1); }
  void unref() {
    if (!_ptr || !_st) return;
    const int left = _st->refs.fetch_sub(1) - 1;
    // every consumer is done: whatever a GPU node left in the device mirror is history now
    if (1 == left && _st->managed) sdrg_buffer_invalidate_device(_ptr);
    if (1 == left && _st->owner) _st->owner->bufferUnused(*this);
    if (0 == left) {
      detail::Storage::release(_ptr, _st->managed);
      delete _st;
      _ptr = 0; _st = 0;
    }
  }
  inline int refCount() const { return _st ? _st->refs.load() : 0; }
  inline bool isUnused() const { return !_st || 1 == _st->refs.load(); }
  /** True if the storage has a device mirror (GPU nodes then skip their own staging). */
  inline bool isDeviceBacked() const { return _st && _st->managed; }

protected:
  char *_ptr;
  size_t _storage_size, _b_offset, _b_length;
  detail::Storage *_st;
};


template <class T>
class Buffer : public RawBuffer {
public:
  Buffer() : RawBuffer(), _size(0) {}
  Buffer(T *data, size_t size) : RawBuffer((char *)data, 0, sizeof(T) * size), _size(size) {}
  Buffer(size_t N, BufferOwner *owner = 0, bool device_backed = false) : RawBuffer(N * sizeof(T), owner, device_backed), _size(N) {}
  Buffer(const Buffer<T> &o) : RawBuffer(o), _size(o._size) {}
  /** Reinterprets the bytes of any buffer as elements of T. */
  explicit Buffer(const RawBuffer &o) : RawBuffer(o), _size(o.bytesLen() / sizeof(T)) {}
  virtual ~Buffer() { _size = 0; }

  const Buffer<T> &operator=(const Buffer<T> o) { RawBuffer::operator=(o); _size = o._size; return *this; }
  inline bool operator<(const Buffer<T> &o) const { return this->_ptr < o._ptr; }

  inline size_t size() const { return _size; }
  inline T &operator[](int idx) const {
#ifdef SDR_DEBUG
    if ((idx < 0) || ((size_t)idx >= _size)) {
      RuntimeError err; err << "Index " << idx << " out of bounds [0," << _size << ")"; throw err;
    }
#endif
    return reinterpret_cast<T *>(_ptr + _b_offset)[idx];
  }

  inline double norm2() const {
    double s = 0;
    for (size_t i = 0; i < _size; i++) s += std::real(std::conj((*this)[i]) * (*this)[i]);
    return std::sqrt(s);
  }
  inline double norm() const {
    double s = 0;
    for (size_t i = 0; i < _size; i++) s += std::abs((*this)[i]);
    return s;
  }
  inline double norm(double p) const {
    double s = 0;
    for (size_t i = 0; i < _size; i++) s += std::pow(std::abs((*this)[i]), p);
    return std::pow(s, 1. / p);
  }
  inline Buffer<T> &operator*=(const T &a) { for (size_t i = 0; i < _size; i++) (*this)[i] *= a; return *this; }
  inline Buffer<T> &operator/=(const T &a) { for (size_t i = 0; i < _size; i++) (*this)[i] /= a; return *this; }

  template <class oT> Buffer<oT> as() const { return Buffer<oT>((const RawBuffer &)(*this)); }

  inline Buffer<T> sub(size_t offset, size_t len) const {
    if ((offset + len) > _size) return Buffer<T>();
    return Buffer<T>(RawBuffer(*this, offset * sizeof(T), len * sizeof(T)));
  }
  inline Buffer<T> head(size_t n) const { return (n > _size) ? Buffer<T>() : sub(0, n); }
  inline Buffer<T> tail(size_t n) const { return (n > _size) ? Buffer<T>() : sub(_size - n, n); }

protected:
  size_t _size;
};

template <class Scalar>
std::ostream &operator<<(std::ostream &stream, const sdr::Buffer<Scalar> &b) {
  stream << "[";
  const size_t n = b.size();
  for (size_t i = 0; i < n; i++) {
    if (n > 10 && i == 5) { stream << ", ..."; i = n - 5; }
    if (i) stream << ", ";
    stream << +b[i];
  }
  return stream << "]";
}


/** A pool of equally sized buffers; buffers return to the pool when the last outside reference is
 * dropped (RawBuffer::unref -> bufferUnused).  Unlike src/buffer.hh:332-342, resize() also makes
 * the new buffers available (the reference forgets to, which breaks BufferNode/FilterNode). */
template <class Scalar>
class BufferSet : public BufferOwner {
public:
  BufferSet(size_t N, size_t size) : _bufferSize(size) { _free.reserve(N); grow(N); }
  virtual ~BufferSet() {
    for (typename std::map<void *, Buffer<Scalar> >::iterator it = _all.begin(); it != _all.end(); ++it) it->second.unref();
  }
  inline bool hasBuffer() { return !_free.empty(); }
  inline Buffer<Scalar> getBuffer() {
    if (_free.empty()) { RuntimeError err; err << "BufferSet: no free buffer"; throw err; }
    void *id = _free.back(); _free.pop_back();
    return _all[id];
  }
  virtual void bufferUnused(const RawBuffer &buffer) {
    if (_all.count(buffer.ptr())) _free.push_back(buffer.ptr());
  }
  void resize(size_t numBuffers) { if (_all.size() < numBuffers) grow(numBuffers - _all.size()); }
protected:
  void grow(size_t n) {
    for (size_t i = 0; i < n; i++) {
      Buffer<Scalar> b(_bufferSize, this);
      _all[b.ptr()] = b; _free.push_back(b.ptr());
    }
  }
  size_t _bufferSize;
  std::map<void *, Buffer<Scalar> > _all;
  std::vector<void *> _free;
};


/** Byte FIFO on top of a RawBuffer (src/buffer.hh:356-470). */
class RawRingBuffer : public RawBuffer {
public:
  RawRingBuffer() : RawBuffer(), _take_idx(0), _b_stored(0) {}
  RawRingBuffer(size_t size) : RawBuffer(size), _take_idx(0), _b_stored(0) {}
  RawRingBuffer(const RawRingBuffer &o) : RawBuffer(o), _take_idx(o._take_idx), _b_stored(o._b_stored) {}
  virtual ~RawRingBuffer() {}
  const RawRingBuffer &operator=(const RawRingBuffer &o) {
    RawBuffer::operator=(o); _take_idx = o._take_idx; _b_stored = o._b_stored; return *this;
  }
  char &operator[](int idx) { return *(ptr() + wrap(_take_idx + (size_t)idx)); }
  inline size_t bytesLen() const { return _b_stored; }
  inline size_t bytesFree() const { return _storage_size - _b_stored; }
  inline bool put(const RawBuffer &src) {
    const size_t n = src.bytesLen();
    if (n > bytesFree()) return false;
    copyIn(wrap(_take_idx + _b_stored), src.data(), n);
    _b_stored += n;
    return true;
  }
  inline bool take(const RawBuffer &dest, size_t N) {
    if (N > dest.bytesLen() || N > _b_stored) return false;
    copyOut(dest.data(), _take_idx, N);
    _take_idx = wrap(_take_idx + N); _b_stored -= N;
    return true;
  }
  inline void drop(size_t N) { if (N > _b_stored) N = _b_stored; _take_idx = wrap(_take_idx + N); _b_stored -= N; }
  inline void clear() { _take_idx = _b_stored = 0; }
  inline void resize(size_t N) {
    if (_storage_size == N) return;
    _take_idx = _b_stored = 0;
    RawBuffer::operator=(RawBuffer(N));
  }
protected:
  inline size_t wrap(size_t i) const { return (_storage_size && i >= _storage_size) ? i - _storage_size : i; }
  void copyIn(size_t at, const char *src, size_t n) {
    const size_t first = (at + n <= _storage_size) ? n : _storage_size - at;
    std::memcpy(_ptr + at, src, first);
    if (first < n) std::memcpy(_ptr, src + first, n - first);
  }
  void copyOut(char *dst, size_t at, size_t n) const {
    const size_t first = (at + n <= _storage_size) ? n : _storage_size - at;
    std::memcpy(dst, _ptr + at, first);
    if (first < n) std::memcpy(dst + first, _ptr, n - first);
  }
  size_t _take_idx, _b_stored;
};

template <class Scalar>
class RingBuffer : public RawRingBuffer {
public:
  RingBuffer() : RawRingBuffer(), _size(0), _stored(0) {}
  RingBuffer(size_t N) : RawRingBuffer(N * sizeof(Scalar)), _size(N), _stored(0) {}
  RingBuffer(const RingBuffer<Scalar> &o) : RawRingBuffer(o), _size(o._size), _stored(o._stored) {}
  virtual ~RingBuffer() {}
  const RingBuffer<Scalar> &operator=(const RingBuffer<Scalar> &o) {
    RawRingBuffer::operator=(o); _size = o._size; _stored = o._stored; return *this;
  }
  Scalar &operator[](int idx) { return reinterpret_cast<Scalar &>(RawRingBuffer::operator[](idx * sizeof(Scalar))); }
  inline size_t stored() const { return _stored; }
  inline size_t free() const { return _size - _stored; }
  inline size_t size() const { return _size; }
  inline bool put(const Buffer<Scalar> &d) { if (!RawRingBuffer::put(d)) return false; _stored += d.size(); return true; }
  inline bool take(const Buffer<Scalar> &d, size_t N) {
    if (!RawRingBuffer::take(d, N * sizeof(Scalar))) return false;
    _stored -= N; return true;
  }
  inline void drop(size_t N) { RawRingBuffer::drop(N * sizeof(Scalar)); _stored = _b_stored / sizeof(Scalar); }
  inline void resize(size_t N) { RawRingBuffer::resize(N * sizeof(Scalar)); _size = N; _stored = 0; }
protected:
  size_t _size, _stored;
};

}  // namespace sdr
#endif
