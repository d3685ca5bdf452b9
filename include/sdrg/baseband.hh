// sdrg/baseband.hh -- IQBaseBand<Scalar> as a GPU node behind libsdr's node interface.
// Same constructors, setters, config()/process() meaning and error behaviour as
// src/baseband.hh:21-297; the work is done by libsdrg (sdrg_iqbb_*).  Scalar: int8_t, int16_t
// (bit-exact w.r.t. the reference) or float (defined in DESIGN.md).
#ifndef SDRG_BASEBAND_HH
#define SDRG_BASEBAND_HH

#include "gpu.hh"
#include "logger.hh"
#include "node.hh"
#include "traits.hh"

namespace sdr {

template <class Scalar>
class IQBaseBand : public Sink< std::complex<Scalar> >, public Source {
public:
  typedef std::complex<Scalar> CScalar;

  /** Filter centre frequency equals the shift frequency Fc (src/baseband.hh:35). */
  IQBaseBand(double Fc, double width, size_t order, size_t sub_sample, double oFs = 0.0) : _h(0) {
    gpu::check(sdrg_iqbb_create(Traits<Scalar>::scalarId, Fc, Fc, width, order, sub_sample, oFs, &_h));
  }
  IQBaseBand(double Fc, double Ff, double width, size_t order, size_t sub_sample, double oFs = 0.0) : _h(0) {
    gpu::check(sdrg_iqbb_create(Traits<Scalar>::scalarId, Fc, Ff, width, order, sub_sample, oFs, &_h));
  }
  virtual ~IQBaseBand() { sdrg_iqbb_destroy(_h); _buffer.unref(); }

  size_t order() const { return info().order; }
  void setOrder(size_t o) { gpu::check(sdrg_iqbb_set_order(_h, o)); }
  void setCenterFrequency(double Fc) { gpu::check(sdrg_iqbb_set_center_frequency(_h, Fc)); }
  void setFilterFrequency(double Ff) { gpu::check(sdrg_iqbb_set_filter_frequency(_h, Ff)); }
  void setFilterWidth(double width) { gpu::check(sdrg_iqbb_set_filter_width(_h, width)); }
  size_t subSample() const { return info().sub_sample; }
  void setSubsample(size_t ss) { gpu::check(sdrg_iqbb_set_subsample(_h, ss)); republish(); }
  void setOutputSampleRate(double Fs) { gpu::check(sdrg_iqbb_set_output_sample_rate(_h, Fs)); republish(); }
  /** int16_t only: consume complex uint8 / int8 buffers directly, AutoCast fused into the load
   * (call before the node is configured; handleBuffer() then reinterprets the raw bytes). */
  void setInputType(Config::Type type) { gpu::check(sdrg_iqbb_set_input_type(_h, (int)type)); _raw8 = (type == Config::Type_cu8 || type == Config::Type_cs8); }
  virtual void handleBuffer(const RawBuffer &buffer, bool allow_overwrite) {
    if (!_raw8) { Sink<CScalar>::handleBuffer(buffer, allow_overwrite); return; }
    if (!_buffer.isUnused()) return;
    run_raw(buffer, buffer.bytesLen() / 2);
  }

  virtual bool acceptsDeviceBuffers() const { return true; }

  virtual void config(const Config &src_cfg) {
    const sdrg_config in = src_cfg.c();
    sdrg_config out;
    gpu::check(sdrg_iqbb_configure(_h, &in, &out));   // throws ConfigError on a type mismatch
    if (SDRG_T_UNDEFINED == out.type) return;         // incomplete config: ignored (baseband.hh:118)
    _src = src_cfg; _out = out;
    _buffer.unref();
    _buffer = Buffer<CScalar>(out.buffer_size, 0, true);
    LogMessage msg(LOG_DEBUG);
    msg << "Configured IQBaseBand node (B200):" << std::endl
        << " type " << src_cfg.type() << std::endl << " sample-rate " << src_cfg.sampleRate() << "Hz" << std::endl
        << " in buffer size " << src_cfg.bufferSize() << std::endl << " sub-sample by " << info().sub_sample << std::endl
        << " out buffer size " << out.buffer_size;
    Logger::get().log(msg);
    this->setConfig(Config::from(out));
  }

  virtual void process(const Buffer<CScalar> &buffer, bool allow_overwrite) {
    if (allow_overwrite) run(buffer, buffer);
    else if (_buffer.isUnused()) run(buffer, _buffer);
    // else: output buffer still in use -> the input is dropped, like the reference (baseband.hh:141-150)
  }

protected:
  sdrg_iqbb_info info() const { sdrg_iqbb_info i; gpu::check(sdrg_iqbb_get_info(_h, &i, 0, 0)); return i; }
  void republish() {    // the rate setters reconfigure and publish a new output config (baseband.hh:156-194)
    if (!_src.hasType() || !_src.hasSampleRate() || !_src.hasBufferSize()) return;
    config(_src);
  }
  void run_raw(const RawBuffer &in, size_t n_in) {       // fused AutoCast: always out of place (2-byte samples in)
    void *st = gpu::stream();
    const void *d_in = gpu::deviceInput(in, st);
    size_t n_out = 0;
    gpu::check(sdrg_iqbb_outputs_for(_h, n_in, &n_out));
    void *d_out = gpu::deviceOutput(_buffer);
    if (d_out) {
      gpu::check(sdrg_iqbb_process_dev(_h, d_in, n_in, d_out, _buffer.size(), &n_out, st));
      if (n_out) gpu::publish(_buffer, n_out * sizeof(CScalar), st);
    } else {                                  // no CUDA-backed storage (should not happen on a GPU box)
      void *d_tmp = 0;
      gpu::check(sdrg_scratch((n_out + 1) * sizeof(CScalar), &d_tmp));
      gpu::check(sdrg_iqbb_process_dev(_h, d_in, n_in, d_tmp, _buffer.size(), &n_out, st));
      gpu::check(sdrg_memcpy_d2h_async(_buffer.data(), d_tmp, n_out * sizeof(CScalar), st));
      gpu::check(sdrg_stream_synchronize(st));
    }
    this->send(_buffer.head(n_out), true);
  }
  void run(const Buffer<CScalar> &in, const Buffer<CScalar> &out) {
    void *st = gpu::stream();
    const void *d_in = gpu::deviceInput(in, st);
    size_t n_out = 0;
    gpu::check(sdrg_iqbb_outputs_for(_h, in.size(), &n_out));
    if (n_out > out.size()) { RuntimeError err; err << "IQBaseBand: output buffer too small"; throw err; }
    void *d_out = gpu::deviceOutput(out);
    const bool staged = (0 == d_out);           // `out` wraps foreign memory: scratch + copy back
    if (staged && n_out) gpu::check(sdrg_scratch_out(n_out * sizeof(CScalar), &d_out));   // thread-private, reused across calls
    // The finalize kernel writes the outputs only after the accumulate kernel has consumed every
    // input sample (stream order), so `out` may alias `in` on the device as it does on the host.
    gpu::check(sdrg_iqbb_process_dev(_h, d_in, in.size(), d_out, n_out, &n_out, st));
    if (staged) {
      if (n_out) { gpu::check(sdrg_memcpy_d2h_async(out.data(), d_out, n_out * sizeof(CScalar), st)); gpu::check(sdrg_stream_synchronize(st)); }
    } else if (n_out) {
      gpu::publish(out, n_out * sizeof(CScalar), st);
    }
    this->send(out.head(n_out), true);
  }

  sdrg_iqbb *_h;
  bool _raw8 = false;
  Config _src;
  sdrg_config _out;
  Buffer<CScalar> _buffer;
};


/** Real-input BaseBand<Scalar> (src/baseband.hh:304-529): complex band-pass FIR on a real stream, NCO,
 * averaging decimator; always out of place into the node's own buffer (baseband.hh:407-418).  Only
 * Scalar = int16_t is built by libsdrg (sdrg_iqbb_create_real); bit-exact w.r.t. the reference. */
template <class Scalar>
class BaseBand : public Sink<Scalar>, public Source {
public:
  typedef std::complex<Scalar> CScalar;

  BaseBand(double Fc, double width, size_t order, size_t sub_sample) : _h(0) {            // Ff = Fc, baseband.hh:322
    gpu::check(sdrg_iqbb_create_real(Traits<Scalar>::scalarId, Fc, Fc, width, order, sub_sample, &_h));
  }
  BaseBand(double Fc, double Ff, double width, size_t order, size_t sub_sample) : _h(0) {
    gpu::check(sdrg_iqbb_create_real(Traits<Scalar>::scalarId, Fc, Ff, width, order, sub_sample, &_h));
  }
  virtual ~BaseBand() { sdrg_iqbb_destroy(_h); _buffer.unref(); }

  /** FreqShiftBase interface (src/freqshift.hh:44-65); sampleRate() is Source's, the OUTPUT rate (the
   * reference's BaseBand inherits two sampleRate() members and cannot be asked without qualification). */
  double inputSampleRate() const { return _src.sampleRate(); }
  void setFrequencyShift(double F) { gpu::check(sdrg_iqbb_set_center_frequency(_h, F)); }
  virtual bool acceptsDeviceBuffers() const { return true; }

  virtual void config(const Config &src_cfg) {
    const sdrg_config in = src_cfg.c();
    sdrg_config out;
    gpu::check(sdrg_iqbb_configure(_h, &in, &out));   // ConfigError "Can not configure BaseBand: Invalid type ..."
    if (SDRG_T_UNDEFINED == out.type) return;         // incomplete config: ignored (baseband.hh:361)
    _src = src_cfg;
    _buffer.unref();
    _buffer = Buffer<CScalar>(out.buffer_size, 0, true);
    LogMessage msg(LOG_DEBUG);
    msg << "Configured BaseBand node (B200):" << std::endl
        << " sample-rate " << src_cfg.sampleRate() << "Hz" << std::endl
        << " in buffer size " << src_cfg.bufferSize() << std::endl << " out buffer size " << out.buffer_size;
    Logger::get().log(msg);
    this->setConfig(Config::from(out));
  }

  virtual void process(const Buffer<Scalar> &buffer, bool /*allow_overwrite*/) {
    if (!_buffer.isUnused()) return;                  // dropped, like the reference (baseband.hh:409-417)
    void *st = gpu::stream();
    const void *d_in = gpu::deviceInput(buffer, st);
    size_t n_out = 0;
    void *d_out = gpu::deviceOutput(_buffer);
    if (!d_out) { RuntimeError err; err << "BaseBand: output buffer has no device storage"; throw err; }
    gpu::check(sdrg_iqbb_process_dev(_h, d_in, buffer.size(), d_out, _buffer.size(), &n_out, st));
    if (n_out) gpu::publish(_buffer, n_out * sizeof(CScalar), st);
    this->send(_buffer.head(n_out), true);
  }

protected:
  sdrg_iqbb *_h;
  Config _src;
  Buffer<CScalar> _buffer;
};

}  // namespace sdr
#endif
