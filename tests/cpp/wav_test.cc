// WavSink<T> / WavSource conformance (include/sdrg/wavfile.hh) against the reference's own files:
//   wav_test <u8|s16|cu8|cs16> <in.raw> <buffer_size> <Fs> <ref.wav> <tmp prefix>
// (1) writes in.raw through WavSink<T>            -> <prefix>.wav   (the driver compares it with the
//     file the reference's WavSink wrote, byte for byte);
// (2) reads ref.wav (written by the REFERENCE) with WavSource -> <prefix>.data/.counts/.cfg, compared
//     with what the reference's WavSource delivered;
// (3) error behaviour (src/wavfile.hh:21-25,48-52,70-76; src/wavfile.cc:51-56).  CPU only.
#include "sdrg/sdr.hh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace sdr;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

static std::vector<char> slurp(const std::string &p) {
  std::vector<char> d; FILE *f = std::fopen(p.c_str(), "rb");
  if (!f) { std::printf("cannot open %s\n", p.c_str()); std::exit(2); }
  char buf[65536]; size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) d.insert(d.end(), buf, buf + n);
  std::fclose(f);
  return d;
}
static void spill(const std::string &p, const void *d, size_t n) {
  FILE *f = std::fopen(p.c_str(), "wb"); std::fwrite(d, 1, n, f); std::fclose(f);
}

template <class T> class Feed : public Source {
public:
  void setup(double Fs, size_t bs) { setConfig(Config(Config::typeId<T>(), Fs, bs, 1)); }
  void push(const Buffer<T> &b) { send(b, false); }
};

class RawCapture : public SinkBase {
public:
  std::vector<char> data; std::vector<uint32_t> sizes; Config cfg;
  virtual void config(const Config &c) { cfg = c; }
  virtual void handleBuffer(const RawBuffer &b, bool) { data.insert(data.end(), b.data(), b.data() + b.bytesLen()); sizes.push_back((uint32_t)b.bytesLen()); }
};

struct EosFlag { int n = 0; void hit() { n++; } };

template <class T>
static void roundtrip(const std::string &in, size_t bs, double Fs, const std::string &ref_wav, const std::string &prefix) {
  const std::vector<char> raw = slurp(in);
  const size_t total = raw.size() / sizeof(T);
  {
    Feed<T> feed; WavSink<T> sink(prefix + ".wav");
    feed.connect(&sink, true); feed.setup(Fs, bs);
    Buffer<T> work(bs);
    for (size_t off = 0; off < total; off += bs) {
      const size_t n = std::min(bs, total - off);
      std::memcpy(work.data(), raw.data() + off * sizeof(T), n * sizeof(T));
      feed.push(work.head(n));
    }
    sink.close();
    sink.close();                                    // idempotent
    feed.push(work.head(1));                         // after close(): ignored (wavfile.hh:113)
  }
  WavSource src(ref_wav, bs); RawCapture cap; EosFlag eos;
  CHECK(src.isOpen());
  CHECK(src.isReal() == (Config::typeId<T>() == Config::Type_u8 || Config::typeId<T>() == Config::Type_s16));
  CHECK(src.frameCount() == total);
  src.addEOS(&eos, &EosFlag::hit);
  src.connect(&cap, true);
  int guard = 0;
  while (src.isOpen() && guard++ < 100000) src.next();
  CHECK(1 == eos.n && !src.isOpen());
  spill(prefix + ".data", cap.data.data(), cap.data.size());
  spill(prefix + ".counts", cap.sizes.data(), cap.sizes.size() * sizeof(uint32_t));
  const double cfg[3] = { (double)cap.cfg.type(), cap.cfg.sampleRate(), (double)cap.cfg.bufferSize() };
  spill(prefix + ".cfg", cfg, sizeof(cfg));
  CHECK(cap.data.size() == raw.size() && 0 == std::memcmp(cap.data.data(), raw.data(), raw.size()));
}

static void error_cases(const std::string &prefix) {
  bool threw = false;
  try { WavSink<float> bad(prefix + ".bad.wav"); } catch (ConfigError &e) { threw = true; CHECK(std::string(e.what()).find("integer typed") != std::string::npos); }
  CHECK(threw);
  threw = false;
  try { WavSink<int16_t> bad("/nonexistent-dir/x.wav"); } catch (ConfigError &) { threw = true; }
  CHECK(threw);
  {
    WavSink<int16_t> sink(prefix + ".cfg.wav"); threw = false;
    try { sink.config(Config(Config::Type_cs16, 48e3, 1024, 1)); } catch (ConfigError &) { threw = true; }
    CHECK(threw);
    sink.config(Config(Config::Type_s16, 0, 1024, 1));                 // incomplete: ignored
  }
  spill(prefix + ".junk", "this is not a wav file at all........................", 48);
  threw = false;
  try { WavSource s(prefix + ".junk"); } catch (RuntimeError &e) { threw = true; CHECK(std::string(e.what()).find("is not a WAV file") != std::string::npos); }
  CHECK(threw);
  {   // a header with an extra chunk before "data", and one without any data chunk
    unsigned char h[64]; std::memset(h, 0, sizeof(h));
    std::memcpy(h, "RIFF", 4); std::memcpy(h + 8, "WAVE", 4); std::memcpy(h + 12, "fmt ", 4);
    h[16] = 16; h[20] = 1; h[22] = 1; h[24] = 0x40; h[25] = 0x1f; h[32] = 2; h[34] = 16;     // PCM, mono, 8000 Hz, 16 bit
    std::memcpy(h + 36, "LIST", 4); h[40] = 4;                                            // 4-byte LIST chunk
    std::memcpy(h + 48, "data", 4); h[52] = 8;                                            // 4 frames
    const int16_t pay[4] = { 1, -2, 300, -32768 }; std::memcpy(h + 56, pay, 8);
    spill(prefix + ".list.wav", h, 64);
    WavSource s(prefix + ".list.wav", 16); RawCapture cap; s.connect(&cap, true);
    CHECK(s.isOpen() && s.frameCount() == 4 && cap.cfg.type() == Config::Type_s16 && cap.cfg.sampleRate() == 8000.0);
    s.next();
    CHECK(cap.data.size() == 8 && 0 == std::memcmp(cap.data.data(), pay, 8));
    spill(prefix + ".nodata.wav", h, 48);
    threw = false;
    try { WavSource t(prefix + ".nodata.wav"); } catch (RuntimeError &e) { threw = true; CHECK(std::string(e.what()).find("no 'data' chunk") != std::string::npos); }
    CHECK(threw);
  }
  WavSource none("/nonexistent-dir/none.wav");       // unopenable: no throw, just not open (wavfile.cc:37)
  CHECK(!none.isOpen());
}

int main(int argc, char **argv) {
  if (argc != 7) { std::printf("usage: wav_test type in bs Fs ref.wav prefix\n"); return 2; }
  const std::string t = argv[1], in = argv[2], ref = argv[5], prefix = argv[6];
  const size_t bs = std::strtoull(argv[3], 0, 10); const double Fs = std::atof(argv[4]);
  if (t == "u8") roundtrip<uint8_t>(in, bs, Fs, ref, prefix);
  else if (t == "s16") roundtrip<int16_t>(in, bs, Fs, ref, prefix);
  else if (t == "cu8") roundtrip< std::complex<uint8_t> >(in, bs, Fs, ref, prefix);
  else if (t == "cs16") roundtrip< std::complex<int16_t> >(in, bs, Fs, ref, prefix);
  else return 2;
  error_cases(prefix);
  if (failures) { std::printf("wav_test: %d failure(s)\n", failures); return 1; }
  std::printf("wav_test: ok\n");
  return 0;
}
