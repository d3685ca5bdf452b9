// sdrg/demod.hh -- AMDemod, USBDemod and FMDemod as GPU nodes behind libsdr's node interface
// (src/demod.hh:16-266).  Same template parameters, config()/process() meaning, in-place rules and
// error behaviour; the arithmetic runs in libsdrg (sdrg_*demod_*).
#ifndef SDRG_DEMOD_HH
#define SDRG_DEMOD_HH

#include "gpu.hh"
#include "logger.hh"
#include "node.hh"
#include "traits.hh"

namespace sdr {

namespace detail {
/** Shared plumbing of the three demodulators: pick the output storage, run, publish, send. */
template <class In, class Out>
class DemodNode : public Sink<In>, public Source {
public:
  virtual ~DemodNode() { _buffer.unref(); }
  virtual bool acceptsDeviceBuffers() const { return true; }
protected:
  /** kernel(d_in, n, d_out, stream, aliased) */
  template <class Kernel>
  void run(const Buffer<In> &in, bool in_place, Kernel kernel, bool overwrite_downstream) {
    const size_t n = in.size();
    Buffer<Out> out = in_place ? Buffer<Out>((const RawBuffer &)in) : _buffer;
    void *st = gpu::stream();
    const void *d_in = gpu::deviceInput(in, st);
    void *d_out = in_place ? const_cast<void *>(d_in) : gpu::deviceOutput(out);
    const bool mirrored = in_place ? (0 != gpu::deviceOutput(in)) : (0 != d_out);
    if (!in_place && !d_out) gpu::check(sdrg_scratch_out(n * sizeof(Out) + 1, &d_out));   // thread-private, reused across calls
    kernel(d_in, n, d_out, st);
    if (mirrored) {
      gpu::publish(out, n * sizeof(Out), st);
    } else {          // foreign memory: bring the result back now
      gpu::check(sdrg_memcpy_d2h_async(out.data(), d_out, n * sizeof(Out), st));
      gpu::check(sdrg_stream_synchronize(st));
    }
    this->send(out.head(n), overwrite_downstream);
  }
  Buffer<Out> _buffer;
};
}  // namespace detail


/** AM demodulator (src/demod.hh:16-86). */
template <class Scalar>
class AMDemod : public detail::DemodNode< std::complex<Scalar>, Scalar > {
public:
  AMDemod() {}
  virtual void config(const Config &src_cfg) {
    const sdrg_config in = src_cfg.c(); sdrg_config out;
    gpu::check(sdrg_amdemod_configure(Traits<Scalar>::scalarId, &in, &out));
    if (SDRG_T_UNDEFINED == out.type) return;
    this->_buffer.unref();
    this->_buffer = Buffer<Scalar>(src_cfg.bufferSize(), 0, true);
    this->setConfig(Config::from(out));
  }
  virtual void process(const Buffer< std::complex<Scalar> > &buffer, bool allow_overwrite) {
    this->run(buffer, allow_overwrite, [](const void *i, size_t n, void *o, void *st) {
      gpu::check(sdrg_amdemod_process_dev(Traits<Scalar>::scalarId, i, n, o, st)); }, true);
  }
};


/** SSB upper-side-band demodulator (src/demod.hh:91-166). */
template <class Scalar>
class USBDemod : public detail::DemodNode< std::complex<Scalar>, Scalar > {
public:
  USBDemod() {}
  virtual void config(const Config &src_cfg) {
    const sdrg_config in = src_cfg.c(); sdrg_config out;
    gpu::check(sdrg_usbdemod_configure(Traits<Scalar>::scalarId, &in, &out));
    if (SDRG_T_UNDEFINED == out.type) return;
    this->_buffer.unref();
    this->_buffer = Buffer<Scalar>(src_cfg.bufferSize(), 0, true);
    this->setConfig(Config::from(out));
  }
  virtual void process(const Buffer< std::complex<Scalar> > &buffer, bool allow_overwrite) {
    this->run(buffer, allow_overwrite, [](const void *i, size_t n, void *o, void *st) {
      gpu::check(sdrg_usbdemod_process_dev(Traits<Scalar>::scalarId, i, n, o, st)); }, false);
  }
};


/** FM demodulator (src/demod.hh:172-266): integer input -> int16 output, float -> float.
 * Element 0 of every buffer is skipped exactly like the reference. */
template <class iScalar, class oScalar = iScalar>
class FMDemod : public detail::DemodNode< std::complex<iScalar>, oScalar > {
public:
  FMDemod() : _h(0), _can_overwrite(false) { gpu::check(sdrg_fmdemod_create(Traits<iScalar>::scalarId, &_h)); }
  virtual ~FMDemod() { sdrg_fmdemod_destroy(_h); }
  virtual void config(const Config &src_cfg) {
    const sdrg_config in = src_cfg.c(); sdrg_config out;
    gpu::check(sdrg_fmdemod_configure(_h, &in, &out));
    if (SDRG_T_UNDEFINED == out.type) return;
    if ((int)Config::typeId<oScalar>() != out.type) {
      ConfigError err; err << "FMDemod: output type " << Config::typeId<oScalar>() << " is not available for input "
                           << src_cfg.type() << " (the device path produces " << (Config::Type)out.type << ")";
      throw err;
    }
    this->_buffer.unref();
    this->_buffer = Buffer<oScalar>(src_cfg.bufferSize(), 0, true);
    _can_overwrite = (sizeof(std::complex<iScalar>) >= sizeof(oScalar));
    this->setConfig(Config::from(out));
  }
  virtual void process(const Buffer< std::complex<iScalar> > &buffer, bool allow_overwrite) {
    if (0 == buffer.size()) return;
    const bool in_place = allow_overwrite && _can_overwrite;
    sdrg_fmdemod *h = _h;
    this->run(buffer, in_place, [h, in_place](const void *i, size_t n, void *o, void *st) {
      gpu::check(sdrg_fmdemod_process_dev(h, i, n, o, in_place ? 1 : 0, st)); }, false);
  }
protected:
  sdrg_fmdemod *_h;
  bool _can_overwrite;
};


/** FM de-emphasis (src/demod.hh:271-362): integer 1-pole IIR with rounding, 75 us time constant. */
template <class Scalar>
class FMDeemph : public Sink<Scalar>, public Source {
public:
  FMDeemph(bool enabled = true) : Sink<Scalar>(), Source(), _enabled(enabled), _h(0) {
    static_assert(sizeof(Scalar) == 2, "the device FMDeemph is implemented for int16_t");
    gpu::check(sdrg_fmdeemph_create(1, &_h));
  }
  virtual ~FMDeemph() { sdrg_fmdeemph_destroy(_h); _buffer.unref(); }
  inline bool isEnabled() const { return _enabled; }
  inline void enable(bool enabled) { _enabled = enabled; }
  virtual bool acceptsDeviceBuffers() const { return true; }
  virtual void config(const Config &src_cfg) {
    const sdrg_config in = src_cfg.c(); sdrg_config out;
    gpu::check(sdrg_fmdeemph_configure(_h, &in, &out));
    if (SDRG_T_UNDEFINED == out.type) return;
    _buffer.unref();
    _buffer = Buffer<Scalar>(src_cfg.bufferSize(), 0, true);
    this->setConfig(Config::from(out));
  }
  virtual void process(const Buffer<Scalar> &buffer, bool allow_overwrite) {
    if (!_enabled) { this->send(buffer, allow_overwrite); return; }
    const Buffer<Scalar> out = allow_overwrite ? buffer : _buffer;
    const size_t n = buffer.size();
    void *st = gpu::stream();
    const void *d_in = gpu::deviceInput(buffer, st);
    void *d_out = allow_overwrite ? const_cast<void *>(d_in) : gpu::deviceOutput(out);
    const bool mirrored = 0 != gpu::deviceOutput(out);
    void *d_tmp = 0;
    if (!d_out) { gpu::check(sdrg_scratch(n * sizeof(Scalar), &d_tmp)); d_out = d_tmp; }
    gpu::check(sdrg_fmdeemph_process_dev(_h, d_in, n, n, d_out, st));    // one thread per stream: in place is safe
    if (mirrored) gpu::publish(out, n * sizeof(Scalar), st);
    else { gpu::check(sdrg_memcpy_d2h_async(out.data(), d_out, n * sizeof(Scalar), st)); gpu::check(sdrg_stream_synchronize(st)); }
    this->send(out.head(n), allow_overwrite);
  }
protected:
  bool _enabled;
  sdrg_fmdeemph *_h;
  Buffer<Scalar> _buffer;
};

}  // namespace sdr
#endif
