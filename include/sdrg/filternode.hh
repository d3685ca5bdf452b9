// sdrg/filternode.hh -- FilterNode<float> / FilterSource<float>: the FFT-convolution filter bank of
// src/filternode.hh:231-283 as a GPU node.  One forward FFT per block is shared by all filters; each
// added filter is a Source of its own (connect sinks to the pointer addFilter() returns; setFreq()
// retunes it).  The input may arrive in buffers of any size: re-chunking to `block_size` samples
// (the reference's BufferNode, which crashes as shipped) happens on the device.
#ifndef SDRG_FILTERNODE_HH
#define SDRG_FILTERNODE_HH

#include <list>
#include <vector>

#include "gpu.hh"
#include "logger.hh"
#include "node.hh"

namespace sdr {

template <class Scalar> class FilterNode;

template <class Scalar>
class FilterSource : public Source {
public:
  /** Retunes the pass band [fmin, fmax] (src/filternode.hh:128-130). */
  void setFreq(double fmin, double fmax) { gpu::check(sdrg_filter_set_freq(_h, _index, fmin, fmax)); }
  size_t index() const { return _index; }
protected:
  friend class FilterNode<Scalar>;
  FilterSource(sdrg_filter *h, size_t index) : Source(), _h(h), _index(index) {}
  sdrg_filter *_h;
  size_t _index;
};

template <class Scalar>
class FilterNode {
  typedef std::complex<Scalar> C;
  class Input : public Sink<C> {
  public:
    Input(FilterNode *n) : _n(n) {}
    virtual bool acceptsDeviceBuffers() const { return true; }
    virtual void config(const Config &cfg) { _n->configure(cfg); }
    virtual void process(const Buffer<C> &b, bool) { _n->run(b); }
  private:
    FilterNode *_n;
  };

public:
  FilterNode(size_t block_size = 1024) : _block(block_size), _h(0), _input(this), _cap(0) {
    static_assert(sizeof(Scalar) == sizeof(float), "the device filter bank is implemented for float");
    gpu::check(sdrg_filter_create(block_size, &_h));
  }
  virtual ~FilterNode() {
    for (typename std::list<FilterSource<Scalar> *>::iterator it = _filters.begin(); it != _filters.end(); ++it) delete *it;
    _out.unref();
    sdrg_filter_destroy(_h);
  }
  /** The sink to connect the upstream source to (src/filternode.hh:257-259). */
  Sink<C> *sink() { return &_input; }
  /** Adds a band-pass [fmin, fmax] to the bank (src/filternode.hh:262-270). */
  FilterSource<Scalar> *addFilter(double fmin, double fmax) {
    size_t idx = 0;
    gpu::check(sdrg_filter_add(_h, fmin, fmax, &idx));
    _filters.push_back(new FilterSource<Scalar>(_h, idx));
    if (_cfg.hasType()) _filters.back()->setConfig(Config(Config::Type_cf32, _cfg.sampleRate(), _block, _cfg.numBuffers()));
    return _filters.back();
  }

protected:
  void configure(const Config &cfg) {
    const sdrg_config in = cfg.c(); sdrg_config out;
    gpu::check(sdrg_filter_configure(_h, &in, &out));
    if (SDRG_T_UNDEFINED == out.type) return;
    _cfg = cfg;
    // room for everything one input buffer can release: (pending + bufferSize) rounded down to blocks
    _cap = ((cfg.bufferSize() + _block - 1) / _block + 1) * _block;
    _out.unref();
    _out = Buffer<C>(_cap * (_filters.empty() ? 1 : _filters.size()), 0, true);
    for (typename std::list<FilterSource<Scalar> *>::iterator it = _filters.begin(); it != _filters.end(); ++it)
      (*it)->setConfig(Config(Config::Type_cf32, cfg.sampleRate(), _block, cfg.numBuffers()));
  }
  void run(const Buffer<C> &in) {
    if (_filters.empty()) return;
    if (_out.size() < _cap * _filters.size()) { _out.unref(); _out = Buffer<C>(_cap * _filters.size(), 0, true); }
    if (!_out.isUnused()) return;                 // downstream still holds the previous output: drop (like the reference's nodes)
    void *st = gpu::stream();
    const void *d_in = gpu::deviceInput(in, st);
    void *d_out = gpu::deviceOutput(_out);
    size_t n_out = 0;
    if (d_out) {
      gpu::check(sdrg_filter_process_dev(_h, d_in, in.size(), d_out, _cap, &n_out, st));
      if (n_out) gpu::publish(_out, _out.bytesLen(), st);
    } else {                                      // no device mirror (tiny buffers): host entry point
      gpu::check(sdrg_stream_synchronize(st));
      gpu::check(sdrg_filter_process(_h, in.data(), in.size(), _out.data(), _cap, &n_out));
    }
    size_t f = 0;
    for (typename std::list<FilterSource<Scalar> *>::iterator it = _filters.begin(); it != _filters.end(); ++it, ++f)
      for (size_t off = 0; off < n_out; off += _block)        // one block per send(), like FilterSource::process
        (*it)->send(_out.sub(f * _cap + off, _block), false);
  }

  size_t _block;
  sdrg_filter *_h;
  Input _input;
  Config _cfg;
  size_t _cap;
  Buffer<C> _out;
  std::list<FilterSource<Scalar> *> _filters;
};

}  // namespace sdr
#endif
