// sdrg/autocast.hh -- AutoCast< std::complex<int16_t> > for complex 8-bit input as a GPU node
// (src/autocast.hh:13-262; the casts used in front of IQBaseBand in examples/sdr_fm.cc:39,49-50 and
// sdr_rec.cc).  cu8: bytes read through int8_t*, (v-127)<<8 (reference quirk, autocast.hh:187-194);
// cs8: v<<8.  Any other cast throws ConfigError with the reference's message.
// When the next node is an IQBaseBand<int16_t>, IQBaseBand::setInputType() fuses this cast into its
// load instead (sdrg_iqbb_set_input_type): 2 bytes per sample from HBM and no extra pass.
#ifndef SDRG_AUTOCAST_HH
#define SDRG_AUTOCAST_HH

#include "gpu.hh"
#include "node.hh"
#include "traits.hh"

namespace sdr {

template <class Scalar>
class AutoCast : public SinkBase, public Source {
public:
  AutoCast() : SinkBase(), Source(), _in_type(SDRG_T_UNDEFINED) {}
  virtual ~AutoCast() { _buffer.unref(); }
  virtual bool acceptsDeviceBuffers() const { return true; }

  virtual void config(const Config &src_cfg) {
    if (!src_cfg.hasType() || !src_cfg.hasBufferSize()) return;
    const int out_type = Traits<Scalar>::scalarId;
    if ((int)src_cfg.type() == out_type) { _in_type = out_type; }            // identity
    else if (out_type == SDRG_T_CS16 && (src_cfg.type() == Config::Type_cu8 || src_cfg.type() == Config::Type_cs8)) {
      _in_type = (int)src_cfg.type();
    } else {
      ConfigError err;
      err << "AutoCast: Can not cast from type " << src_cfg.type() << " to " << (Config::Type)out_type;
      throw err;
    }
    _buffer.unref();
    _buffer = Buffer<Scalar>(src_cfg.bufferSize(), 0, true);
    this->setConfig(Config((Config::Type)out_type, src_cfg.sampleRate(), src_cfg.bufferSize(), 1));
  }

  virtual void handleBuffer(const RawBuffer &buffer, bool allow_overwrite) {
    if (_in_type == Traits<Scalar>::scalarId) { this->send(buffer, allow_overwrite); return; }
    if (!_buffer.isUnused()) return;                       // output still in use: drop (autocast.hh:100-108)
    const size_t n = buffer.bytesLen() / 2;                // complex 8-bit samples
    void *st = gpu::stream();
    const void *d_in = gpu::deviceInput(buffer, st);
    void *d_out = gpu::deviceOutput(_buffer);
    if (d_out) {
      gpu::check(sdrg_autocast_process_dev(_in_type, SDRG_T_CS16, d_in, n, d_out, st));
      gpu::publish(_buffer, n * sizeof(Scalar), st);
    } else {
      gpu::check(sdrg_stream_synchronize(st));
      gpu::check(sdrg_autocast_process(_in_type, SDRG_T_CS16, buffer.data(), n, _buffer.data()));
    }
    this->send(_buffer.head(n), true);
  }

protected:
  int _in_type;
  Buffer<Scalar> _buffer;
};

}  // namespace sdr
#endif
