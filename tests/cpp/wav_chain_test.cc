// File-to-file receive chain on the GPU nodes, the shape of cmd/ and examples/sdr_* tools working on
// recorded audio-band IF (src/wavfile.hh, src/baseband.hh:304-529):
//   WavSource (real int16) --direct--> BaseBand<int16_t> --direct--> FMDemod<int16_t> --direct--> WavSink<int16_t>
// checked bit for bit against the oracle fed the same buffers.  Needs a GPU.
#include "wavfile.hh"      // reference header names through include/sdrg/compat
#include "baseband.hh"
#include "demod.hh"
#include "../../oracle/sdr_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace sdr;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

class Feed : public Source {
public:
  void setup(double Fs, size_t bs) { setConfig(Config(Config::Type_s16, Fs, bs, 1)); }
  void push(const Buffer<int16_t> &b) { send(b, false); }
};

int main(int argc, char **argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp";
  const double Fs = 96e3; const size_t bs = 4096, N = 10 * bs + 777, ss = 4;
  std::vector<int16_t> x(N);
  uint32_t lcg = 99;
  for (size_t n = 0; n < N; n++) {
    const double t = n / Fs;
    lcg = lcg * 1664525u + 1013904223u;
    x[n] = int16_t(14000 * cos(2 * M_PI * 12e3 * t + 4 * sin(2 * M_PI * 440 * t)) + 3000 * cos(2 * M_PI * 31e3 * t) + int((lcg >> 16) % 65) - 32);
  }
  {   // record the IF
    Feed feed; WavSink<int16_t> rec(dir + "/if.wav");
    feed.connect(&rec, true); feed.setup(Fs, bs);
    Buffer<int16_t> work(bs);
    for (size_t off = 0; off < N; off += bs) {
      const size_t n = std::min(bs, N - off);
      std::memcpy(work.data(), &x[off], n * sizeof(int16_t));
      feed.push(work.head(n));
    }
  }
  // oracle on the same buffer cuts
  std::vector<int16_t> want;
  {
    orc_iqbb *s = new orc_iqbb; int16_t last = 0;
    orc_rbb_init(s, 12e3, 12e3, 8e3, 31, ss); orc_rbb_config(s, Fs, bs);
    std::vector<int16_t> bb(2 * (bs + 2)), fm(bs + 2);
    for (size_t off = 0; off < N; off += bs) {
      const size_t n = std::min(bs, N - off);
      const size_t m = orc_rbb_process(s, &x[off], n, bb.data());
      if (!m) continue;
      fm[0] = bb[0];                                   // in place: element 0 shows in[0].real()
      orc_fmdemod_s16(bb.data(), m, fm.data(), &last);
      want.insert(want.end(), fm.begin(), fm.begin() + m);
    }
    delete s;
  }
  {   // the chain
    WavSource src(dir + "/if.wav", bs);
    CHECK(src.isOpen() && src.isReal() && src.sampleRate() == Fs);
    BaseBand<int16_t> baseband(12e3, 8e3, 31, ss);
    FMDemod<int16_t> demod;
    WavSink<int16_t> audio(dir + "/audio.wav");
    src.connect(&baseband, true); baseband.connect(&demod, true); demod.connect(&audio, true);
    CHECK(baseband.sampleRate() == Fs / ss && demod.sampleRate() == Fs / ss);
    CHECK(baseband.type() == Config::Type_cs16 && demod.type() == Config::Type_s16);
    bool threw = false;                                // complex input is a type error (baseband.hh:363-369)
    try { BaseBand<int16_t> b2(1e3, 1e3, 9, 2); b2.config(Config(Config::Type_cs16, Fs, bs, 1)); } catch (ConfigError &) { threw = true; }
    CHECK(threw);
    int guard = 0;
    while (src.isOpen() && guard++ < 1000) src.next();
    audio.close();
  }
  // compare the audio file's payload
  FILE *f = std::fopen((dir + "/audio.wav").c_str(), "rb");
  CHECK(f != 0);
  if (f) {
    unsigned char h[44]; CHECK(std::fread(h, 1, 44, f) == 44);
    CHECK(0 == std::memcmp(h, "RIFF", 4) && 0 == std::memcmp(h + 36, "data", 4));
    const uint32_t rate = h[24] | (h[25] << 8) | (h[26] << 16) | ((uint32_t)h[27] << 24);
    CHECK(rate == (uint32_t)(Fs / ss));
    std::vector<int16_t> got(want.size() + 16);
    const size_t n = std::fread(got.data(), sizeof(int16_t), got.size(), f);
    std::fclose(f);
    CHECK(n == want.size());
    CHECK(n == N / ss);
    size_t bad = 0;
    for (size_t i = 0; i < std::min(n, want.size()); i++) bad += got[i] != want[i];
    CHECK(0 == bad);
    if (bad) std::printf("%zu of %zu audio samples differ\n", bad, n);
  }
  if (failures) { std::printf("wav_chain_test: %d failure(s)\n", failures); return 1; }
  std::printf("wav_chain_test: ok\n");
  return 0;
}
