// FilterNode<float> / FFTPlan<float> through the C++ node classes against the oracle's
// FilterSink+FilterSource restatement (tolerance 1e-5 relative RMS).  Needs a GPU.
#include "sdrg/sdr.hh"
#include "../../oracle/sdr_oracle.h"

#include <cmath>
#include <cstdio>
#include <vector>

using namespace sdr;
typedef std::complex<float> cf;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

static double rel_rms(const std::vector<cf> &a, const std::vector<cf> &b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < b.size(); i++) { num += std::norm(std::complex<double>(a[i]) - std::complex<double>(b[i])); den += std::norm(std::complex<double>(b[i])); }
  return std::sqrt(num / (den > 0 ? den : 1));
}

class Feed : public Source {
public:
  void setup(double Fs, size_t bs) { setConfig(Config(Config::Type_cf32, Fs, bs, 1)); }
  void push(const Buffer<cf> &b) { send(b, false); }
};
class Capture : public Sink<cf> {
public:
  std::vector<cf> data; size_t calls = 0; Config cfg;
  virtual void config(const Config &c) { cfg = c; }
  virtual void process(const Buffer<cf> &b, bool) { for (size_t i = 0; i < b.size(); i++) data.push_back(b[i]); ++calls; }
};

int main() {
  const size_t N = 1024, nblk = 24; const double Fs = 2.4e6;
  std::vector<cf> x(N * nblk);
  uint32_t lcg = 7;
  for (size_t n = 0; n < x.size(); n++) {
    const double t = n / Fs;
    lcg = lcg * 1664525u + 1013904223u; const double nr = ((lcg >> 8) % 2001) / 1e5 - 0.01;
    x[n] = cf(float(0.5 * cos(2 * M_PI * 150e3 * t) + 0.3 * cos(2 * M_PI * -400e3 * t + 1) + nr),
              float(0.5 * sin(2 * M_PI * 150e3 * t) + 0.3 * sin(2 * M_PI * -400e3 * t + 1)));
  }
  // oracle
  std::vector<cf> ref1(x.size()), ref2(x.size());
  {
    std::vector<float> k1(4 * N), k2(4 * N), l1(2 * N, 0.f), l2(2 * N, 0.f);
    orc_filter_design_f32(N, 100e3, 200e3, Fs, k1.data()); orc_filter_design_f32(N, -500e3, -300e3, Fs, k2.data());
    for (size_t b = 0; b < nblk; b++) {
      orc_filter_ola_block_f32(N, k1.data(), (const float *)&x[b * N], (float *)&ref1[b * N], l1.data());
      orc_filter_ola_block_f32(N, k2.data(), (const float *)&x[b * N], (float *)&ref2[b * N], l2.data());
    }
  }
  {   // the bank as a node: input in buffers of 1536 samples (not a multiple of the block size)
    Feed feed; FilterNode<float> bank(N);
    FilterSource<float> *f1 = bank.addFilter(100e3, 200e3), *f2 = bank.addFilter(-300e3, -500e3);   // swapped bounds are fixed up
    Capture c1, c2;
    f1->connect(&c1, true); f2->connect(&c2, true);
    feed.connect(bank.sink(), true);
    feed.setup(Fs, 1536);
    CHECK(c1.cfg.bufferSize() == N); CHECK(c1.cfg.type() == Config::Type_cf32);
    Buffer<cf> work(1536);
    for (size_t off = 0; off < x.size(); off += 1536) {
      memcpy(work.data(), &x[off], 1536 * sizeof(cf));
      feed.push(work);
    }
    CHECK(c1.data.size() == x.size()); CHECK(c1.calls == nblk);
    CHECK(rel_rms(c1.data, ref1) < 1e-5); CHECK(rel_rms(c2.data, ref2) < 1e-5);
    work.unref();
  }
  {   // FFTPlan<float>: forward then backward returns n * x; an empty buffer throws ConfigError
    const size_t n = 4096;
    Buffer<cf> a(n), b(n);
    for (size_t i = 0; i < n; i++) a[i] = x[i];
    FFTPlan<float> fwd(a, b, FFT::FORWARD); fwd();
    std::vector<float> oin(2 * n), oout(2 * n);
    memcpy(oin.data(), a.data(), n * sizeof(cf)); orc_fft_f32(oin.data(), oout.data(), n, +1);
    std::vector<cf> got(n), want(n);
    for (size_t i = 0; i < n; i++) { got[i] = b[i]; want[i] = cf(oout[2 * i], oout[2 * i + 1]); }
    CHECK(rel_rms(got, want) < 2e-6);
    FFT::exec(b, FFT::BACKWARD);
    for (size_t i = 0; i < n; i++) { got[i] = b[i] / float(n); want[i] = a[i]; }
    CHECK(rel_rms(got, want) < 2e-6);
    bool threw = false;
    try { Buffer<cf> e; FFTPlan<float> p(e, FFT::FORWARD); } catch (ConfigError &) { threw = true; }
    CHECK(threw);
    {   // any size, like the reference's FFTW plan: 96 points (Bluestein), forward then backward returns n * x
      Buffer<cf> o(96), O(96), back(96);
      for (size_t i = 0; i < 96; i++) o[i] = cf(float(i % 7) - 3.0f, float(i % 5) - 2.0f);
      FFT::exec(o, O, FFT::FORWARD); FFT::exec(O, back, FFT::BACKWARD);
      double err = 0, ref = 0;
      for (size_t i = 0; i < 96; i++) { err += std::norm(std::complex<double>(back[i]) / 96.0 - std::complex<double>(o[i])); ref += std::norm(std::complex<double>(o[i])); }
      CHECK(std::sqrt(err / ref) < 1e-5);
      std::complex<double> dc(0, 0);
      for (size_t i = 0; i < 96; i++) dc += std::complex<double>(o[i]);
      CHECK(std::abs(std::complex<double>(O[0]) - dc) < 1e-3);
    }
    {   // FFTPlan<double>
      Buffer< std::complex<double> > o(100), O(100);
      for (size_t i = 0; i < 100; i++) o[i] = std::complex<double>(double(i % 9) - 4.0, double(i % 4) - 1.5);
      FFT::exec(o, O, FFT::FORWARD);
      std::complex<double> dc(0, 0), nyq(0, 0);
      for (size_t i = 0; i < 100; i++) { dc += o[i]; nyq += (i % 2 ? -1.0 : 1.0) * o[i]; }
      CHECK(std::abs(O[0] - dc) < 1e-9); CHECK(std::abs(O[50] - nyq) < 1e-9);
    }
    a.unref(); b.unref();
  }
  std::printf(failures ? "filter_test: %d FAILED\n" : "filter_test: ok\n", failures);
  return failures ? 1 : 0;
}
