// sdrg/autocast.hh -- AutoCast<Scalar> as a GPU node (src/autocast.hh:13-262): Scalar = int8_t, complex<int8_t>,
// int16_t or complex<int16_t> from every source type of the reference's table (autocast.hh:30-69), each cast with the
// reference's own arithmetic (e.g. cu8 -> cs16, the cast in front of IQBaseBand in examples/sdr_fm.cc:39,49-50: bytes
// read through int8_t*, (v-127)<<8).  A pair outside the table throws ConfigError with the reference's message.
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
    if (!src_cfg.hasType() || !src_cfg.hasSampleRate() || !src_cfg.hasBufferSize()) return;       // autocast.hh:27
    const int out_type = Traits<Scalar>::scalarId;
    size_t probe = 0;
    if (SDRG_OK != sdrg_autocast_out_bytes((int)src_cfg.type(), out_type, 1, &probe)) {
      ConfigError err;
      err << "AutoCast: Can not cast from type " << src_cfg.type() << " to " << (Config::Type)out_type;
      throw err;
    }
    _in_type = (int)src_cfg.type();
    static const size_t elem[] = {0, 1, 1, 2, 2, 4, 8, 2, 2, 4, 4, 8, 16};
    _in_elem = elem[_in_type];
    _buffer.unref();
    _buffer = Buffer<Scalar>(src_cfg.bufferSize(), 0, true);                                        // autocast.hh:79
    this->setConfig(Config((Config::Type)out_type, src_cfg.sampleRate(), src_cfg.bufferSize(), 1));
  }

  virtual void handleBuffer(const RawBuffer &buffer, bool allow_overwrite) {
    if (SDRG_T_UNDEFINED == _in_type) return;              // no conversion selected (autocast.hh:96)
    if (_in_type == Traits<Scalar>::scalarId) { this->send(buffer, allow_overwrite); return; }      // identity: forwarded
    if (!_buffer.isUnused()) return;                       // output still in use: drop
    const size_t n = buffer.bytesLen() / _in_elem;         // input elements
    size_t out_bytes = 0;
    gpu::check(sdrg_autocast_out_bytes(_in_type, Traits<Scalar>::scalarId, n, &out_bytes));
    void *st = gpu::stream();
    const void *d_in = gpu::deviceInput(buffer, st);
    void *d_out = gpu::deviceOutput(_buffer);
    if (d_out) {
      gpu::check(sdrg_autocast_process_dev(_in_type, Traits<Scalar>::scalarId, d_in, n, d_out, st));
      gpu::publish(_buffer, out_bytes, st);
    } else {
      gpu::check(sdrg_stream_synchronize(st));
      gpu::check(sdrg_autocast_process(_in_type, Traits<Scalar>::scalarId, buffer.data(), n, _buffer.data()));
    }
    this->send(RawBuffer(_buffer, 0, out_bytes), false);   // autocast.hh:103
  }

protected:
  int _in_type;
  size_t _in_elem = 1;
  Buffer<Scalar> _buffer;
};

}  // namespace sdr
#endif
