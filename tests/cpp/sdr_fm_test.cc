// The complete examples/sdr_fm.cc chain minus the hardware ends, with libsdr's own header names
// (compat include dir):  cu8 source -> AutoCast<cs16> -> IQBaseBand<int16_t>(100e3, 12.5e3, 21, 1, 8000)
// -> FMDemod<int16_t> -> FMDeemph<int16_t> -> capture, same connect() pattern (direct / queued) as
// examples/sdr_fm.cc:48-53, checked bit for bit against the oracle.  A second run fuses the AutoCast
// into IQBaseBand's load.  Needs a GPU.
#include "demod.hh"
#include "baseband.hh"
#include "autocast.hh"
#include "../../oracle/sdr_oracle.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <vector>

using namespace sdr;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)
typedef std::complex<uint8_t> cu8;

class RtlLike : public Source {     // stands in for RTLSource (cu8 @ 1 MS/s, rtlsource.cc:133-145)
public:
  void setup(double Fs, size_t bs) { setConfig(Config(Config::Type_cu8, Fs, bs, 15)); }
  void push(const RawBuffer &b) { send(b, false); }
};
class Audio : public Sink<int16_t> {
public:
  std::vector<int16_t> data; std::vector<size_t> sizes;
  virtual void config(const Config &) {}
  virtual void process(const Buffer<int16_t> &b, bool) { for (size_t i = 0; i < b.size(); i++) data.push_back(b[i]); sizes.push_back(b.size()); }
};

int main() {
  const double Fs = 1e6; const size_t bs = 131072 / 8, nbuf = 8;
  std::vector<uint8_t> x(2 * bs * nbuf);
  uint32_t lcg = 99;
  for (size_t n = 0; n < bs * nbuf; n++) {
    const double t = n / Fs, ph = 2 * M_PI * 100e3 * t + 3.0 * sin(2 * M_PI * 1e3 * t);   // FM, 1 kHz tone
    lcg = lcg * 1664525u + 1013904223u;
    x[2 * n] = uint8_t(127 + 100 * cos(ph) + int((lcg >> 16) % 5) - 2);
    x[2 * n + 1] = uint8_t(127 + 100 * sin(ph) + int((lcg >> 24) % 5) - 2);
  }
  // oracle
  std::vector<int16_t> want; std::vector<size_t> want_sizes;
  {
    orc_iqbb *s = new orc_iqbb; int16_t last = 0, avg = 0;
    orc_iqbb_init(s, ORC_S16, 100e3, 100e3, 12.5e3, 21, 1, 8000.0);
    orc_iqbb_set_center_frequency(s, 100e3); orc_iqbb_set_filter_frequency(s, 100e3);
    orc_iqbb_config(s, Fs, bs);
    const int alpha = orc_fmdeemph_alpha(8000.0);
    std::vector<int16_t> c(2 * bs), bb(2 * (bs + 2)), fm(bs + 2), de(bs + 2);
    for (size_t b = 0; b < nbuf; b++) {
      orc_autocast_u8_s16(&x[2 * b * bs], 2 * bs, c.data());
      const size_t n = orc_iqbb_process(s, c.data(), bs, bb.data());
      fm[0] = bb[0];                                   // FMDemod in place: element 0 = in[0].real()
      orc_fmdemod_s16(bb.data(), n, fm.data(), &last);
      orc_fmdeemph_s16(fm.data(), n, de.data(), alpha, &avg);
      want.insert(want.end(), de.begin(), de.begin() + n); want_sizes.push_back(n);
    }
    delete s;
  }
  for (int fused = 0; fused < 2; fused++) {
    RtlLike src;
    AutoCast< std::complex<int16_t> > cast;
    IQBaseBand<int16_t> baseband(100e3, 12.5e3, 21, 1, 8000.0);
    baseband.setCenterFrequency(100e3);
    baseband.setFilterFrequency(100e3);
    FMDemod<int16_t> demod;
    FMDeemph<int16_t> deemph;
    Audio audio;
    if (fused) { baseband.setInputType(Config::Type_cu8); src.connect(&baseband); }
    else { src.connect(&cast, true); cast.connect(&baseband); }       // sdr_fm.cc:49-50
    baseband.connect(&demod, true);
    demod.connect(&deemph, true);
    deemph.connect(&audio);                                            // queued, like the PortSink
    src.setup(Fs, bs);
    Queue::get().start();
    std::vector< Buffer<cu8> > bufs;                   // one buffer per transfer, like a ring that never wraps
    for (size_t b = 0; b < nbuf; b++) {
      bufs.push_back(Buffer<cu8>(bs));
      memcpy(bufs.back().data(), &x[2 * b * bs], 2 * bs);
      src.push(bufs.back());
      // the queued hop between cast and baseband (sdr_fm.cc:50) drops input while the previous buffer
      // is still in flight, exactly like the reference; pace the source so that nothing is dropped
      for (int spin = 0; spin < 20000 && audio.sizes.size() < b + 1 && Queue::get().isRunning(); spin++)
        std::this_thread::sleep_for(std::chrono::microseconds(100));
    }
    Queue::get().stop(); Queue::get().wait();
    for (size_t b = 0; b < bufs.size(); b++) bufs[b].unref();
    CHECK(audio.sizes == want_sizes);
    CHECK(audio.data == want);
  }
  std::printf(failures ? "sdr_fm_test: %d FAILED\n" : "sdr_fm_test: ok\n", failures);
  return failures ? 1 : 0;
}
