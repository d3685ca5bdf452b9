// Drop-in acceptance test: a chain shaped like examples/sdr_fm.cc:38-53 / sdr_rec.cc:66-109 built
// from the sdr:: node classes of include/sdrg/ (GPU nodes) and checked, bit for bit, against the
// oracle (oracle/sdr_oracle.c) fed the same buffers.  Needs a GPU.
//   source --direct--> IQBaseBand<int16_t> --direct--> FMDemod<int16_t> --direct--> capture     (in place)
//   source --queued--> IQBaseBand<int16_t> --direct--> {FMDemod, AMDemod, USBDemod} -> captures (out of place)
#include "sdrg/sdr.hh"
#include "../../oracle/sdr_oracle.h"

#include <cmath>
#include <cstdio>
#include <vector>

using namespace sdr;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

typedef std::complex<int16_t> cs16;

static void synth(std::vector<cs16> &x, double Fs) {
  uint32_t lcg = 12345;
  for (size_t n = 0; n < x.size(); n++) {
    const double t = n / Fs;
    double re = 8192 * cos(2 * M_PI * 103e3 * t) + 4096 * cos(2 * M_PI * 99e3 * t + 0.5);
    double im = 8192 * sin(2 * M_PI * 103e3 * t) + 4096 * sin(2 * M_PI * 99e3 * t + 0.5);
    lcg = lcg * 1664525u + 1013904223u; re += int((lcg >> 16) % 129) - 64;
    lcg = lcg * 1664525u + 1013904223u; im += int((lcg >> 16) % 129) - 64;
    x[n] = cs16(int16_t(re), int16_t(im));
  }
}

class Feed : public Source {
public:
  void setup(double Fs, size_t bs) { setConfig(Config(Config::Type_cs16, Fs, bs, 1)); }
  void push(const Buffer<cs16> &b, bool ow) { send(b, ow); }
};

template <class T> class Capture : public Sink<T> {
public:
  std::vector<T> data; std::vector<size_t> sizes; Config cfg;
  virtual void config(const Config &c) { cfg = c; }
  virtual void process(const Buffer<T> &b, bool) {      // a plain host sink: reads through operator[]
    for (size_t i = 0; i < b.size(); i++) data.push_back(b[i]);
    sizes.push_back(b.size());
  }
};

int main() {
  const double Fs = 1e6; const size_t bs = 16384, nbuf = 6;
  std::vector<cs16> x(bs * nbuf); synth(x, Fs);

  // oracle: same construction as examples/sdr_fm.cc:40-42
  std::vector<int16_t> o_bb, o_fm, o_am, o_usb; std::vector<size_t> o_sizes;
  {
    orc_iqbb *s = new orc_iqbb; int16_t last = 0;
    orc_iqbb_init(s, ORC_S16, 100e3, 100e3, 12.5e3, 21, 1, 8000.0);
    orc_iqbb_set_center_frequency(s, 100e3); orc_iqbb_set_filter_frequency(s, 100e3);
    orc_iqbb_config(s, Fs, bs);
    std::vector<int16_t> out(2 * (bs + 2)), fm(bs + 2), am(bs + 2), usb(bs + 2);
    for (size_t b = 0; b < nbuf; b++) {
      const size_t n = orc_iqbb_process(s, &x[b * bs], bs, out.data());
      o_sizes.push_back(n);
      o_bb.insert(o_bb.end(), out.begin(), out.begin() + 2 * n);
      orc_amdemod_s16(out.data(), n, am.data()); orc_usbdemod_s16(out.data(), n, usb.data());
      fm[0] = out[0];                                  // in-place view: element 0 shows in[0].real()
      orc_fmdemod_s16(out.data(), n, fm.data(), &last);
      o_fm.insert(o_fm.end(), fm.begin(), fm.begin() + n);
      o_am.insert(o_am.end(), am.begin(), am.begin() + n); o_usb.insert(o_usb.end(), usb.begin(), usb.begin() + n);
    }
    delete s;
  }

  {   // (1) all direct, in place -- the sdr_fm.cc shape
    Feed feed; IQBaseBand<int16_t> baseband(100e3, 12.5e3, 21, 1, 8000.0);
    baseband.setCenterFrequency(100e3); baseband.setFilterFrequency(100e3);
    FMDemod<int16_t> demod; Capture<int16_t> audio;
    feed.connect(&baseband, true); baseband.connect(&demod, true); demod.connect(&audio, true);
    feed.setup(Fs, bs);
    CHECK(audio.cfg.type() == Config::Type_s16); CHECK(audio.cfg.sampleRate() == 8000.0);
    Buffer<cs16> work(bs);
    for (size_t b = 0; b < nbuf; b++) {
      memcpy(work.data(), &x[b * bs], bs * sizeof(cs16));
      feed.push(work, true);
    }
    CHECK(audio.sizes == o_sizes);
    CHECK(audio.data == o_fm);
    work.unref();
  }
  {   // (2) queued into the base band, three demodulators on one source => all out of place
    Feed feed; IQBaseBand<int16_t> baseband(100e3, 100e3, 12.5e3, 21, 1, 8000.0);
    baseband.setCenterFrequency(100e3); baseband.setFilterFrequency(100e3);
    FMDemod<int16_t> fm; AMDemod<int16_t> am; USBDemod<int16_t> usb;
    Capture<int16_t> c_fm, c_am, c_usb; Capture<cs16> c_bb;
    feed.connect(&baseband, false);
    baseband.connect(&fm, true); baseband.connect(&am, true); baseband.connect(&usb, true); baseband.connect(&c_bb, true);
    fm.connect(&c_fm, true); am.connect(&c_am, true); usb.connect(&c_usb, true);
    feed.setup(Fs, bs);
    Queue::get().start();
    BufferSet<cs16> pool(3, bs);
    for (size_t b = 0; b < nbuf; b++) {
      while (!pool.hasBuffer()) std::this_thread::yield();       // pool buffers come back when the queue unrefs them
      Buffer<cs16> w = pool.getBuffer();
      memcpy(w.data(), &x[b * bs], bs * sizeof(cs16));
      feed.push(w, false);
    }
    Queue::get().stop(); Queue::get().wait();
    CHECK(c_bb.sizes == o_sizes);
    std::vector<int16_t> bb_flat; for (size_t i = 0; i < c_bb.data.size(); i++) { bb_flat.push_back(c_bb.data[i].real()); bb_flat.push_back(c_bb.data[i].imag()); }
    CHECK(bb_flat == o_bb);
    CHECK(c_am.data == o_am); CHECK(c_usb.data == o_usb);
    // out-of-place FM: element 0 of every buffer is never written (demod.hh:245); compare the rest
    CHECK(c_fm.data.size() == o_fm.size());
    size_t off = 0, bad = 0;
    for (size_t b = 0; b < o_sizes.size(); off += o_sizes[b], b++)
      for (size_t i = 1; i < o_sizes[b]; i++) bad += (c_fm.data[off + i] != o_fm[off + i]);
    CHECK(bad == 0);
  }
  {   // (3) error behaviour: type mismatch throws ConfigError at connect time (baseband.hh:120-125)
    Feed feed; feed.setConfig(Config(Config::Type_cf32, Fs, bs, 1));
    IQBaseBand<int16_t> baseband(100e3, 12.5e3, 21, 1, 8000.0);
    bool threw = false;
    try { feed.connect(&baseband, true); } catch (ConfigError &) { threw = true; }
    CHECK(threw);
  }
  {   // (4) small heap-backed buffers (no device mirror) through in-place AM / USB / FM nodes: the input is
      //     staged in scratch memory and the aliased result must not clobber it before it is read
    const size_t n = 300;
    std::vector<int16_t> want_am(n), want_usb(n), want_fm(n);
    orc_amdemod_s16((const int16_t *)&x[0], n, want_am.data()); orc_usbdemod_s16((const int16_t *)&x[0], n, want_usb.data());
    int16_t last = 0; want_fm[0] = x[0].real(); orc_fmdemod_s16((const int16_t *)&x[0], n, want_fm.data(), &last);
    Feed feed; AMDemod<int16_t> am; USBDemod<int16_t> usb; FMDemod<int16_t> fm;
    Capture<int16_t> c_am, c_usb, c_fm;
    Feed f_am, f_usb, f_fm;
    f_am.connect(&am, true); am.connect(&c_am, true);
    f_usb.connect(&usb, true); usb.connect(&c_usb, true);
    f_fm.connect(&fm, true); fm.connect(&c_fm, true);
    f_am.setup(8000.0, n); f_usb.setup(8000.0, n); f_fm.setup(8000.0, n);
    Buffer<cs16> w1(n), w2(n), w3(n);                   // 1200 bytes each: heap storage
    CHECK(!w1.isDeviceBacked());
    memcpy(w1.data(), &x[0], n * sizeof(cs16)); memcpy(w2.data(), &x[0], n * sizeof(cs16)); memcpy(w3.data(), &x[0], n * sizeof(cs16));
    f_am.push(w1, true); f_usb.push(w2, true); f_fm.push(w3, true);
    CHECK(c_am.data == want_am); CHECK(c_usb.data == want_usb); CHECK(c_fm.data == want_fm);
    w1.unref(); w2.unref(); w3.unref();
  }
  std::printf(failures ? "chain_test: %d FAILED\n" : "chain_test: ok\n", failures);
  return failures ? 1 : 0;
}
