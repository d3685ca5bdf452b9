// sdrg/fftplan.hh -- FFTPlan<float> and FFT::exec with the reference's surface
// (src/fftplan.hh:11-34, src/fftplan_fftw3.hh:79-142), executed by libsdrg's FFT kernels instead
// of FFTW3.  Unnormalised; FORWARD = exp(-i..).  Any size 1..2^24 (an empty buffer throws
// ConfigError like the reference); FFTPlan<double> computes in double on the device as well.
#ifndef SDRG_FFTPLAN_HH
#define SDRG_FFTPLAN_HH

#include "buffer.hh"
#include "gpu.hh"

namespace sdr {

class FFT {
public:
  typedef enum { FORWARD, BACKWARD } Direction;
  template <class Scalar>
  static void exec(const Buffer< std::complex<Scalar> > &in, const Buffer< std::complex<Scalar> > &out, FFT::Direction dir);
  template <class Scalar>
  static void exec(const Buffer< std::complex<Scalar> > &inplace, FFT::Direction dir);
};

template <class Scalar> class FFTPlan {};

template <>
class FFTPlan<float> {
public:
  FFTPlan(const Buffer< std::complex<float> > &in, const Buffer< std::complex<float> > &out, FFT::Direction dir)
    : _in(in), _out(out), _h(0) {
    if (in.size() != out.size()) { ConfigError err; err << "Can not construct FFT plan: input & output buffers are of different size!"; throw err; }
    if (in.isEmpty() || out.isEmpty()) { ConfigError err; err << "Can not construct FFT plan: input or output buffer is empty!"; throw err; }
    gpu::check(sdrg_fft_create(in.size(), FFT::BACKWARD == dir ? 1 : 0, &_h));
  }
  FFTPlan(const Buffer< std::complex<float> > &inplace, FFT::Direction dir) : _in(inplace), _out(inplace), _h(0) {
    if (inplace.isEmpty()) { ConfigError err; err << "Can not construct FFT plan: Buffer is empty!"; throw err; }
    gpu::check(sdrg_fft_create(inplace.size(), FFT::BACKWARD == dir ? 1 : 0, &_h));
  }
  virtual ~FFTPlan() { sdrg_fft_destroy(_h); }
  /** Performs the transform on the host buffers given at construction (copies included). */
  void operator()() { gpu::check(sdrg_fft_exec(_h, _in.data(), _out.data(), 1)); }
protected:
  Buffer< std::complex<float> > _in, _out;
  sdrg_fft *_h;
private:
  FFTPlan(const FFTPlan &);
  FFTPlan &operator=(const FFTPlan &);
};

/** FFTPlan<double> (src/fftplan_fftw3.hh:12-75). */
template <>
class FFTPlan<double> {
public:
  FFTPlan(const Buffer< std::complex<double> > &in, const Buffer< std::complex<double> > &out, FFT::Direction dir)
    : _in(in), _out(out), _h(0) {
    if (in.size() != out.size()) { ConfigError err; err << "Can not construct FFT plan: input & output buffers are of different size!"; throw err; }
    if (in.isEmpty() || out.isEmpty()) { ConfigError err; err << "Can not construct FFT plan: input or output buffer is empty!"; throw err; }
    gpu::check(sdrg_fft64_create(in.size(), FFT::BACKWARD == dir ? 1 : 0, &_h));
  }
  FFTPlan(const Buffer< std::complex<double> > &inplace, FFT::Direction dir) : _in(inplace), _out(inplace), _h(0) {
    if (inplace.isEmpty()) { ConfigError err; err << "Can not construct FFT plan: Buffer is empty!"; throw err; }
    gpu::check(sdrg_fft64_create(inplace.size(), FFT::BACKWARD == dir ? 1 : 0, &_h));
  }
  virtual ~FFTPlan() { sdrg_fft64_destroy(_h); }
  void operator()() { gpu::check(sdrg_fft64_exec(_h, _in.data(), _out.data(), 1)); }
protected:
  Buffer< std::complex<double> > _in, _out;
  sdrg_fft64 *_h;
private:
  FFTPlan(const FFTPlan &);
  FFTPlan &operator=(const FFTPlan &);
};

template <class Scalar>
void FFT::exec(const Buffer< std::complex<Scalar> > &in, const Buffer< std::complex<Scalar> > &out, FFT::Direction dir) {
  FFTPlan<Scalar> plan(in, out, dir); plan();
}
template <class Scalar>
void FFT::exec(const Buffer< std::complex<Scalar> > &inplace, FFT::Direction dir) {
  FFTPlan<Scalar> plan(inplace, dir); plan();
}

}  // namespace sdr
#endif
