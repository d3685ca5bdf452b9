// sdrg/node.hh -- Config, SinkBase, Sink<T>, Source, Proxy of the host-side mirror.
// Interface mirrored from src/node.hh:35-332 / src/node.cc:66-114:
//   * Source::connect(sink, direct) stores the link and immediately calls sink->config(config);
//   * Source::setConfig() propagates only on change;
//   * Source::send(): direct links are called synchronously on the sender's thread, queued ones go
//     through Queue; the overwrite permission survives only if there is exactly one sink.
// Addition for device residency: a sink declares with acceptsDeviceBuffers() that it can consume a
// buffer whose current bytes live in the device mirror (GPU nodes do); for every other sink send()
// first brings the bytes back to the host (sdrg_buffer_sync_to_host: a no-op for host data).
// A device-valid mark lives exactly as long as the send() that published it (or, for queued links,
// until the Queue drops its reference): afterwards the host copy is authoritative again, so a CPU
// node refilling a pooled buffer can never be shadowed by stale device bytes.
#ifndef SDRG_NODE_HH
#define SDRG_NODE_HH

#include <complex>
#include <list>
#include <map>
#include <ostream>

#include "buffer.hh"
#include "queue.hh"
#include "exception.hh"

namespace sdr {

class Config {
public:
  typedef enum {
    Type_UNDEFINED = SDRG_T_UNDEFINED, Type_u8 = SDRG_T_U8, Type_s8 = SDRG_T_S8, Type_u16 = SDRG_T_U16,
    Type_s16 = SDRG_T_S16, Type_f32 = SDRG_T_F32, Type_f64 = SDRG_T_F64, Type_cu8 = SDRG_T_CU8,
    Type_cs8 = SDRG_T_CS8, Type_cu16 = SDRG_T_CU16, Type_cs16 = SDRG_T_CS16, Type_cf32 = SDRG_T_CF32,
    Type_cf64 = SDRG_T_CF64
  } Type;

  Config() : _type(Type_UNDEFINED), _sampleRate(0), _bufferSize(0), _numBuffers(0) {}
  Config(Type type, double sampleRate, size_t bufferSize, size_t numBuffers)
    : _type(type), _sampleRate(sampleRate), _bufferSize(bufferSize), _numBuffers(numBuffers) {}
  bool operator==(const Config &o) const {
    return o._type == _type && o._sampleRate == _sampleRate && o._bufferSize == _bufferSize && o._numBuffers == _numBuffers;
  }
  inline bool hasType() const { return Type_UNDEFINED != _type; }
  inline Type type() const { return _type; }
  inline void setType(Type type) { _type = type; }
  inline bool hasSampleRate() const { return 0 != _sampleRate; }
  inline double sampleRate() const { return _sampleRate; }
  inline void setSampleRate(double rate) { _sampleRate = rate; }
  inline bool hasBufferSize() const { return 0 != _bufferSize; }
  inline size_t bufferSize() const { return _bufferSize; }
  inline void setBufferSize(size_t size) { _bufferSize = size; }
  inline bool hasNumBuffers() const { return 0 != _numBuffers; }
  inline size_t numBuffers() const { return _numBuffers; }
  inline void setNumBuffers(size_t N) { _numBuffers = N; }

  template <typename T> static inline Type typeId();

  // C ABI view
  sdrg_config c() const { sdrg_config c; c.type = (int)_type; c.sample_rate = _sampleRate; c.buffer_size = _bufferSize; c.num_buffers = _numBuffers; return c; }
  static Config from(const sdrg_config &c) { return Config((Type)c.type, c.sample_rate, c.buffer_size, c.num_buffers); }

protected:
  Type _type;
  double _sampleRate;
  size_t _bufferSize, _numBuffers;
};

#define SDRG_TYPEID(T, ID) template <> inline Config::Type Config::typeId< T >() { return ID; }
SDRG_TYPEID(uint8_t, Type_u8) SDRG_TYPEID(int8_t, Type_s8) SDRG_TYPEID(uint16_t, Type_u16)
SDRG_TYPEID(int16_t, Type_s16) SDRG_TYPEID(float, Type_f32) SDRG_TYPEID(double, Type_f64)
SDRG_TYPEID(std::complex<uint8_t>, Type_cu8) SDRG_TYPEID(std::complex<int8_t>, Type_cs8)
SDRG_TYPEID(std::complex<uint16_t>, Type_cu16) SDRG_TYPEID(std::complex<int16_t>, Type_cs16)
SDRG_TYPEID(std::complex<float>, Type_cf32) SDRG_TYPEID(std::complex<double>, Type_cf64)
#undef SDRG_TYPEID

inline const char *typeName(Config::Type type) {
  static const char *names[] = {"UNDEFINED", "uint8", "int8", "uint16", "int16", "float", "double", "complex uint8",
                                "complex int8", "complex uint16", "complex int16", "complex float", "complex double"};
  return ((int)type >= 0 && (int)type <= (int)Config::Type_cf64) ? names[(int)type] : "unknown";
}
inline std::ostream &operator<<(std::ostream &stream, Config::Type type) {
  return stream << typeName(type) << " (" << (int)type << ")";
}


class SinkBase {
public:
  SinkBase() {}
  virtual ~SinkBase() {}
  virtual void handleBuffer(const RawBuffer &buffer, bool allow_overwrite) = 0;
  virtual void config(const Config &src_cfg) = 0;
  /** True if the sink reads its input through the device mirror (GPU nodes). */
  virtual bool acceptsDeviceBuffers() const { return false; }
};

template <class Scalar>
class Sink : public SinkBase {
public:
  Sink() : SinkBase() {}
  virtual ~Sink() {}
  virtual void process(const Buffer<Scalar> &buffer, bool allow_overwrite) = 0;
  virtual void handleBuffer(const RawBuffer &buffer, bool allow_overwrite) {
    this->process(Buffer<Scalar>(buffer), allow_overwrite);
  }
};

inline void Queue::deliver(Message &msg) {
  if (!msg.sink()->acceptsDeviceBuffers()) {
    sdrg_buffer_sync_to_host(msg.buffer().data(), msg.buffer().bytesLen());
    if (msg.allowOverwrite()) sdrg_buffer_invalidate_device(msg.buffer().ptr());
  }
  msg.sink()->handleBuffer(msg.buffer(), msg.allowOverwrite());
}


class Source {
public:
  Source() {}
  virtual ~Source() { for (std::list<DelegateInterface *>::iterator it = _eos.begin(); it != _eos.end(); ++it) delete *it; }

  virtual void send(const RawBuffer &buffer, bool allow_overwrite = false) {
    const bool exclusive = allow_overwrite && (1 == _sinks.size());
    for (std::map<SinkBase *, bool>::iterator it = _sinks.begin(); it != _sinks.end(); ++it) {
      if (it->second) {        // direct: same thread, now
        if (!it->first->acceptsDeviceBuffers()) {
          sdrg_buffer_sync_to_host(buffer.data(), buffer.bytesLen());
          if (exclusive) sdrg_buffer_invalidate_device(buffer.ptr());   // the sink may write on the host
        }
        it->first->handleBuffer(buffer, exclusive);
      } else {
        Queue::get().send(buffer, it->first, exclusive);
      }
    }
    if (buffer.isUnused() && buffer.isDeviceBacked()) sdrg_buffer_invalidate_device(buffer.ptr());
  }
  void connect(SinkBase *sink, bool direct = false) { _sinks[sink] = direct; sink->config(_config); }
  void disconnect(SinkBase *sink) { _sinks.erase(sink); }
  virtual void setConfig(const Config &config) {
    if (config == _config) return;
    _config = config;
    propagateConfig(_config);
  }
  virtual double sampleRate() const { return _config.sampleRate(); }
  virtual Config::Type type() const { return _config.type(); }
  template <class T> void addEOS(T *instance, void (T::*function)()) { _eos.push_back(new Delegate<T>(instance, function)); }

protected:
  void signalEOS() { for (std::list<DelegateInterface *>::iterator it = _eos.begin(); it != _eos.end(); ++it) (**it)(); }
  void propagateConfig(const Config &) {
    for (std::map<SinkBase *, bool>::iterator it = _sinks.begin(); it != _sinks.end(); ++it) it->first->config(_config);
  }
  Config _config;
  std::map<SinkBase *, bool> _sinks;
  std::list<DelegateInterface *> _eos;
};


/** Interface of a blocking source (src/node.hh:267-311, src/node.cc:137-190): an input that waits for data from a
 * device or a file.  next() is driven either by the Queue's idle signal (connect_idle) or by the source's own thread
 * (parallel); stop_queue_on_eos wires the EOS signal to Queue::stop; both drivers call next() only while the source
 * is active and the Queue is running.  Deviation: start() raises _is_active.  The reference never sets the flag
 * (node.cc:154-159; no class derives from BlockingSource there), so its thread would leave _parallel_main at once
 * and an idle-driven source would never be polled; a subclass that set it beforehand would make start() return
 * early (`if (_is_active) return`). */
class BlockingSource : public Source {
public:
  BlockingSource(bool parallel = false, bool connect_idle = true, bool stop_queue_on_eos = false)
    : Source(), _is_active(false), _is_parallel(parallel) {
    if (!parallel && connect_idle) Queue::get().addIdle(this, &BlockingSource::_nonvirt_idle_cb);
    if (stop_queue_on_eos) this->addEOS(&Queue::get(), &Queue::stop);
  }
  virtual ~BlockingSource() {
    if (isActive()) stop();
    if (_thread.joinable()) _thread.join();
    Queue::get().remIdle(this);
  }
  virtual void next() = 0;
  inline bool isActive() const { return _is_active; }
  virtual void start() {
    if (_is_active) return;
    _is_active = true;
    if (_is_parallel) {
      if (_thread.joinable()) _thread.join();
      _thread = std::thread(&BlockingSource::_parallel_main, this);
    }
  }
  virtual void stop() {
    if (!_is_active) return;
    _is_active = false;
    if (_is_parallel && _thread.joinable()) _thread.join();
  }

protected:
  void _parallel_main() { while (_is_active && Queue::get().isRunning()) this->next(); }
  void _nonvirt_idle_cb() { if (_is_active && Queue::get().isRunning()) this->next(); }
  volatile bool _is_active;
  bool _is_parallel;
  std::thread _thread;
};


/** Forwards config and buffers unchanged (src/node.hh:312-328). */
class Proxy : public SinkBase, public Source {
public:
  Proxy() : SinkBase(), Source() {}
  virtual ~Proxy() {}
  virtual void config(const Config &src_cfg) { this->setConfig(src_cfg); }
  virtual void handleBuffer(const RawBuffer &buffer, bool) { this->send(buffer); }
};

}  // namespace sdr
#endif
