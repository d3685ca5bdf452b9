// ref_harness.cc -- drives the UNMODIFIED libsdr reference classes on raw binary input.
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile against the reference sources where they
// lie (/root/reference/src, never copied into this repo); the binary goes to oracle/_ref/.
// It (a) generates the golden vectors under tests/golden/ (tests/golden/gen_golden.py),
// (b) validates oracle/sdr_oracle.c, and (c) is the "reference" CPU baseline bench.py times.
//
// Sub-commands (all files are raw little-endian arrays):
//   bb  <s16|s8> <in.bin> <buffer_size> <Fs> <Fc> <Ff> <width> <order> <sub_sample> <oFs> <setcf> <prefix>
//         runs IQBaseBand<T> -> {FMDemod (in place), AMDemod, USBDemod} with direct connections and
//         writes <prefix>.params/.bb/.counts/.fm/.am/.usb
//   ola <in.bin cf32> <block> <Fs> <fmin> <fmax> <prefix>
//         runs FilterSink<float> -> FilterSource<float> with a double-precision FFTPlan stand-in
//         (FFTW3 is not installed); writes <prefix>.kern/.taps/.out
//   rbb <in.bin int16> <buffer_size> <Fs> <Fc> <Ff> <width> <order> <sub_sample> <prefix>
//         runs the real-input BaseBand<int16_t> (src/baseband.hh:304-529); writes <prefix>.params/.bb/.counts
//   wav <u8|s16|cu8|cs16> <in.bin> <buffer_size> <Fs> <out.wav> <prefix>
//         writes the input through WavSink<T> (src/wavfile.hh:16-128), reads it back with WavSource
//         (src/wavfile.cc); writes <prefix>.data/.counts/.cfg (type, rate, buffer size as 3 doubles)
//   cast <cu8|cs8> <in.bin> <buffer_size> <prefix>
//         runs AutoCast< std::complex<int16_t> > on complex 8-bit input; writes <prefix>.cs16
//   deemph <in.bin int16> <buffer_size> <Fs> <prefix>
//         runs FMDeemph<int16_t> (in place and out of place give the same samples); writes <prefix>.out
//   time <s16|s8> <in.bin> <buffer_size> <Fs> <Fc> <Ff> <width> <order> <sub_sample> <oFs> <threads> <min_seconds>
//         times IQBaseBand<T> -> FMDemod, in place, direct connections, T independent chains;
//         prints one JSON line with input Msamples/s
#include "operators.hh"
#include "buffer.hh"
#include "node.hh"
#include "fftplan.hh"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

extern "C" void orc_fft_f32(const float *in, float *out, size_t n, int dir);

// Stand-in for the FFTW3-backed plan (src/fftplan_fftw3.hh:79-142): same surface, double DFT.
namespace sdr {
template <> class FFTPlan<float> {
public:
  FFTPlan(const Buffer< std::complex<float> > &in, const Buffer< std::complex<float> > &out,
          FFT::Direction dir) : _in(in), _out(out), _dir(dir) {
    if (in.size() != out.size()) { ConfigError err; err << "size mismatch"; throw err; }
    if (in.isEmpty() || out.isEmpty()) { ConfigError err; err << "empty"; throw err; }
  }
  FFTPlan(const Buffer< std::complex<float> > &inplace, FFT::Direction dir)
    : _in(inplace), _out(inplace), _dir(dir) {
    if (inplace.isEmpty()) { ConfigError err; err << "empty"; throw err; }
  }
  virtual ~FFTPlan() {}
  void operator() () {
    orc_fft_f32((const float *)_in.data(), (float *)_out.data(), _in.size(),
                (FFT::FORWARD == _dir) ? +1 : -1);
  }
protected:
  Buffer< std::complex<float> > _in, _out;
  FFT::Direction _dir;
};
}

#include "baseband.hh"
#include "demod.hh"
#include "filternode.hh"
#include "autocast.hh"
#include "wavfile.hh"

using namespace sdr;

static std::vector<char> slurp(const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<char> d(n);
  if (n && fread(d.data(), 1, n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
  fclose(f);
  return d;
}

static FILE *wopen(const std::string &p) {
  FILE *f = fopen(p.c_str(), "wb");
  if (!f) { fprintf(stderr, "cannot write %s\n", p.c_str()); exit(2); }
  return f;
}

// A source that hands out the caller's work buffer.
template <class T> class Feed : public Source {
public:
  void setup(double Fs, size_t bs) { setConfig(Config(Config::typeId<T>(), Fs, bs, 1)); }
  void push(const Buffer<T> &b, bool ow) { send(b, ow); }
};

// File-dump sink.
template <class T> class Dump : public Sink<T> {
public:
  Dump() : f(0), total(0) {}
  virtual void config(const Config &) {}
  virtual void process(const Buffer<T> &b, bool) {
    if (f) fwrite(b.data(), sizeof(T), b.size(), f);
    total += b.size(); sizes.push_back((uint32_t)b.size());
  }
  FILE *f; size_t total; std::vector<uint32_t> sizes;
};

// Exposes the protected design results of IQBaseBand.
template <class S> struct BB : public IQBaseBand<S> {
  BB(double Fc, double Ff, double w, size_t o, size_t ss, double oFs) : IQBaseBand<S>(Fc, Ff, w, o, ss, oFs) {}
  void dump(FILE *f) {
    int64_t hdr[4] = { (int64_t)this->_order, (int64_t)this->_sub_sample, (int64_t)this->_lut_inc,
                       (int64_t)(0 > this->_freq_shift) };
    fwrite(hdr, sizeof(hdr), 1, f);
    for (size_t i = 0; i < this->_order; i++) {
      int32_t v[2] = { this->_kernel[i].real(), this->_kernel[i].imag() }; fwrite(v, sizeof(v), 1, f);
    }
    for (size_t i = 0; i < 128; i++) {
      int32_t v[2] = { (int32_t)this->_lut[i].real(), (int32_t)this->_lut[i].imag() }; fwrite(v, sizeof(v), 1, f);
    }
  }
};

// Receives the base-band output, stores it, and re-feeds private copies to the three demods.
template <class S, class FMO> class Tee : public Sink< std::complex<S> > {
public:
  typedef std::complex<S> C;
  Tee(const std::string &prefix) : bs(0) {
    fbb = wopen(prefix + ".bb");
    fm_dump.f = wopen(prefix + ".fm"); am_dump.f = wopen(prefix + ".am"); usb_dump.f = wopen(prefix + ".usb");
    to_fm.connect(&fm, true); fm.connect(&fm_dump, true);
    to_am.connect(&am, true); am.connect(&am_dump, true);
    to_usb.connect(&usb, true); usb.connect(&usb_dump, true);
  }
  virtual void config(const Config &cfg) {
    if (!cfg.hasType() || !cfg.hasBufferSize()) return;
    bs = cfg.bufferSize();
    w_fm = Buffer<C>(bs); w_am = Buffer<C>(bs); w_usb = Buffer<C>(bs);
    to_fm.setConfig(cfg); to_am.setConfig(cfg); to_usb.setConfig(cfg);
  }
  virtual void process(const Buffer<C> &b, bool) {
    fwrite(b.data(), sizeof(C), b.size(), fbb);
    counts.push_back((uint32_t)b.size());
    memcpy(w_fm.data(), b.data(), b.size() * sizeof(C));
    memcpy(w_am.data(), b.data(), b.size() * sizeof(C));
    memcpy(w_usb.data(), b.data(), b.size() * sizeof(C));
    to_fm.push(w_fm.head(b.size()), true);     // single direct sink => in place
    to_am.push(w_am.head(b.size()), true);
    to_usb.push(w_usb.head(b.size()), true);
  }
  void close() { fclose(fbb); fclose(fm_dump.f); fclose(am_dump.f); fclose(usb_dump.f); }
  size_t bs; FILE *fbb; std::vector<uint32_t> counts;
  Buffer<C> w_fm, w_am, w_usb;
  Feed<C> to_fm, to_am, to_usb;
  FMDemod<S, FMO> fm; AMDemod<S> am; USBDemod<S> usb;
  Dump<FMO> fm_dump; Dump<S> am_dump; Dump<S> usb_dump;
};

template <class S, class FMO>
static int run_bb(char **a) {
  typedef std::complex<S> C;
  std::vector<char> raw = slurp(a[0]);
  size_t bs = strtoull(a[1], 0, 10);
  double Fs = atof(a[2]), Fc = atof(a[3]), Ff = atof(a[4]), width = atof(a[5]);
  size_t order = strtoull(a[6], 0, 10), ss = strtoull(a[7], 0, 10);
  double oFs = atof(a[8]); int setcf = atoi(a[9]);
  std::string prefix = a[10];
  size_t total = raw.size() / sizeof(C);

  Feed<C> feed;
  BB<S> bb(Fc, Ff, width, order, ss, oFs);
  if (setcf) { bb.setCenterFrequency(Fc); bb.setFilterFrequency(Ff); }  // as examples/sdr_fm.cc:41-42
  Tee<S, FMO> tee(prefix);
  feed.connect(&bb, true);
  bb.connect(&tee, true);
  feed.setup(Fs, bs);

  FILE *fp = wopen(prefix + ".params"); bb.dump(fp); fclose(fp);

  Buffer<C> work(bs);
  for (size_t off = 0; off < total; off += bs) {
    size_t n = std::min(bs, total - off);
    memcpy(work.data(), raw.data() + off * sizeof(C), n * sizeof(C));
    feed.push(work.head(n), true);    // in place, like a direct single-sink connection
  }
  tee.close();
  FILE *fc = wopen(prefix + ".counts");
  fwrite(tee.counts.data(), sizeof(uint32_t), tee.counts.size(), fc); fclose(fc);
  return 0;
}

// Capture sink for the OLA filter: ref()/unref() lets BufferSet recycle its single buffer.
class OlaDump : public Sink< std::complex<float> > {
public:
  OlaDump(FILE *f_) : f(f_) {}
  virtual void config(const Config &) {}
  virtual void process(const Buffer< std::complex<float> > &b, bool) {
    fwrite(b.data(), sizeof(std::complex<float>), b.size(), f);
    RawBuffer r(b); r.ref(); r.unref();
  }
  FILE *f;
};

struct FS : public FilterSource<float> {
  FS(size_t n, double a, double b) : FilterSource<float>(n, a, b) {}
  void dump(FILE *f) { fwrite(_kern.data(), sizeof(std::complex<float>), _kern.size(), f); }
};

static int run_ola(char **a) {
  typedef std::complex<float> C;
  std::vector<char> raw = slurp(a[0]);
  size_t block = strtoull(a[1], 0, 10);
  double Fs = atof(a[2]), fmin = atof(a[3]), fmax = atof(a[4]);
  std::string prefix = a[5];
  size_t total = raw.size() / sizeof(C);

  Feed<C> feed;
  FilterSink<float> fsink(block);
  FS fsrc(block, fmin, fmax);
  FILE *fo = wopen(prefix + ".out");
  OlaDump dump(fo);
  feed.connect(&fsink, true);
  fsink.connect(&fsrc, true);
  fsrc.connect(&dump, true);
  feed.setup(Fs, block);
  FILE *fk = wopen(prefix + ".kern"); fsrc.dump(fk); fclose(fk);
  FILE *ft = wopen(prefix + ".taps");
  for (size_t i = 0; i < block; i++) {
    C v = sinc_flt_kernel<float>(i, block, std::max(fmin, -Fs/2) + (std::min(fmax, Fs/2) - std::max(fmin, -Fs/2)) / 2,
                                 std::min(fmax, Fs/2) - std::max(fmin, -Fs/2), Fs);
    fwrite(&v, sizeof(C), 1, ft);
  }
  fclose(ft);
  Buffer<C> work(block);
  for (size_t off = 0; off + block <= total; off += block) {
    memcpy(work.data(), raw.data() + off * sizeof(C), block * sizeof(C));
    feed.push(work, false);
  }
  fclose(fo);
  return 0;
}

template <class S> struct RBB : public BaseBand<S> {
  RBB(double Fc, double Ff, double w, size_t o, size_t ss) : BaseBand<S>(Fc, Ff, w, o, ss) {}
  void dump(FILE *f) {
    int64_t hdr[4] = { (int64_t)this->_order, (int64_t)this->_sub_sample, (int64_t)FreqShiftBase<S>::_lut_inc,
                       (int64_t)(0 > FreqShiftBase<S>::_freq_shift) };
    fwrite(hdr, sizeof(hdr), 1, f);
    for (size_t i = 0; i < this->_order; i++) {
      int32_t v[2] = { this->_kernel[i].real(), this->_kernel[i].imag() }; fwrite(v, sizeof(v), 1, f);
    }
  }
};

template <class S>
static int run_rbb(char **a) {
  std::vector<char> raw = slurp(a[0]);
  size_t bs = strtoull(a[1], 0, 10);
  double Fs = atof(a[2]), Fc = atof(a[3]), Ff = atof(a[4]), width = atof(a[5]);
  size_t order = strtoull(a[6], 0, 10), ss = strtoull(a[7], 0, 10);
  std::string prefix = a[8];
  size_t total = raw.size() / sizeof(S);
  Feed<S> feed; RBB<S> bb(Fc, Ff, width, order, ss); Dump< std::complex<S> > dump;
  dump.f = wopen(prefix + ".bb");
  feed.connect(&bb, true); bb.connect(&dump, true);
  feed.setup(Fs, bs);
  FILE *fp = wopen(prefix + ".params"); bb.dump(fp); fclose(fp);
  Buffer<S> work(bs);
  for (size_t off = 0; off < total; off += bs) {
    size_t n = std::min(bs, total - off);
    memcpy(work.data(), raw.data() + off * sizeof(S), n * sizeof(S));
    feed.push(work.head(n), false);
  }
  fclose(dump.f);
  FILE *fc = wopen(prefix + ".counts"); fwrite(dump.sizes.data(), sizeof(uint32_t), dump.sizes.size(), fc); fclose(fc);
  return 0;
}

// Raw capture sink for WavSource (its element type is only known after open()).
class RawDump : public SinkBase {
public:
  RawDump() : f(0) {}
  virtual void config(const Config &c) { cfg = c; }
  virtual void handleBuffer(const RawBuffer &b, bool) {
    fwrite(b.data(), 1, b.bytesLen(), f); sizes.push_back((uint32_t)b.bytesLen());
  }
  FILE *f; Config cfg; std::vector<uint32_t> sizes;
};

template <class T>
static int run_wav(char **a) {
  std::vector<char> raw = slurp(a[0]);
  size_t bs = strtoull(a[1], 0, 10); double Fs = atof(a[2]);
  std::string wav = a[3], prefix = a[4];
  size_t total = raw.size() / sizeof(T);
  {
    Feed<T> feed; WavSink<T> sink(wav);
    feed.connect(&sink, true); feed.setup(Fs, bs);
    Buffer<T> work(bs);
    for (size_t off = 0; off < total; off += bs) {
      size_t n = std::min(bs, total - off);
      memcpy(work.data(), raw.data() + off * sizeof(T), n * sizeof(T));
      feed.push(work.head(n), false);
    }
    sink.close();
  }
  WavSource src(wav, bs); RawDump dump; dump.f = wopen(prefix + ".data");
  src.connect(&dump, true);
  while (src.isOpen()) src.next();
  fclose(dump.f);
  FILE *fc = wopen(prefix + ".counts"); fwrite(dump.sizes.data(), sizeof(uint32_t), dump.sizes.size(), fc); fclose(fc);
  double cfg[3] = { (double)dump.cfg.type(), dump.cfg.sampleRate(), (double)dump.cfg.bufferSize() };
  FILE *fg = wopen(prefix + ".cfg"); fwrite(cfg, sizeof(cfg), 1, fg); fclose(fg);
  return 0;
}

template <class T>
static int run_cast(char **a) {
  std::vector<char> raw = slurp(a[0]);
  size_t bs = strtoull(a[1], 0, 10);
  std::string prefix = a[2];
  size_t total = raw.size() / sizeof(T);
  Feed<T> feed; AutoCast< std::complex<int16_t> > cast; Dump< std::complex<int16_t> > dump;
  dump.f = wopen(prefix + ".cs16");
  feed.connect(&cast, true); cast.connect(&dump, true);
  feed.setup(1e6, bs);
  Buffer<T> work(bs);
  for (size_t off = 0; off < total; off += bs) {
    size_t n = std::min(bs, total - off);
    memcpy(work.data(), raw.data() + off * sizeof(T), n * sizeof(T));
    feed.push(work.head(n), false);
  }
  fclose(dump.f);
  return 0;
}

// AutoCast<Out> fed raw buffers of an arbitrary source type (src/autocast.hh:30-69: the whole cast table).
// args: in_type_id out_type_id in.bin buffer_bytes prefix  ->  prefix.out (bytes), exit 4 when the reference refuses the pair
class RawFeed : public Source {
public:
  void setup(Config::Type t, double Fs, size_t bs) { setConfig(Config(t, Fs, bs, 1)); }
  void push(const RawBuffer &b) { send(b, false); }
};
template <class Out>
static int run_castx_t(Config::Type in_type, size_t in_elem, char **a) {
  std::vector<char> raw = slurp(a[0]);
  size_t bytes_per_buf = strtoull(a[1], 0, 10);
  std::string prefix = a[2];
  RawFeed feed; AutoCast<Out> cast; RawDump dump;
  feed.connect(&cast, true); cast.connect(&dump, true);
  try { feed.setup(in_type, 1e6, bytes_per_buf / in_elem); }
  catch (ConfigError &e) { return 4; }
  dump.f = wopen(prefix + ".out");
  RawBuffer work(bytes_per_buf);
  for (size_t off = 0; off < raw.size(); off += bytes_per_buf) {
    size_t n = std::min(bytes_per_buf, raw.size() - off);
    memcpy(work.data(), raw.data() + off, n);
    feed.push(RawBuffer(work, 0, n));
  }
  fclose(dump.f);
  return 0;
}
static int run_castx(char **a) {
  const int in_t = atoi(a[0]), out_t = atoi(a[1]);
  static const size_t elem[] = {0, 1, 1, 2, 2, 4, 8, 2, 2, 4, 4, 8, 16};
  if (in_t < 1 || in_t > 12) return 2;
  switch (out_t) {
    case Config::Type_s8: return run_castx_t<int8_t>((Config::Type)in_t, elem[in_t], a + 2);
    case Config::Type_cs8: return run_castx_t< std::complex<int8_t> >((Config::Type)in_t, elem[in_t], a + 2);
    case Config::Type_s16: return run_castx_t<int16_t>((Config::Type)in_t, elem[in_t], a + 2);
    case Config::Type_cs16: return run_castx_t< std::complex<int16_t> >((Config::Type)in_t, elem[in_t], a + 2);
  }
  return 2;
}

static int run_deemph(char **a) {
  std::vector<char> raw = slurp(a[0]);
  size_t bs = strtoull(a[1], 0, 10);
  double Fs = atof(a[2]);
  std::string prefix = a[3];
  size_t total = raw.size() / sizeof(int16_t);
  Feed<int16_t> feed; FMDeemph<int16_t> de; Dump<int16_t> dump;
  dump.f = wopen(prefix + ".out");
  feed.connect(&de, true); de.connect(&dump, true);
  feed.setup(Fs, bs);
  Buffer<int16_t> work(bs);
  for (size_t off = 0; off < total; off += bs) {
    size_t n = std::min(bs, total - off);
    memcpy(work.data(), raw.data() + off * sizeof(int16_t), n * sizeof(int16_t));
    feed.push(work.head(n), true);
  }
  fclose(dump.f);
  return 0;
}

template <class T> class Null : public Sink<T> {
public:
  Null() : n(0), acc(0) {}
  virtual void config(const Config &) {}
  virtual void process(const Buffer<T> &b, bool) { n += b.size(); if (b.size() > 1) acc += b[1]; }
  size_t n; long acc;
};

template <class S, class FMO>
static int run_time(char **a) {
  typedef std::complex<S> C;
  std::vector<char> raw = slurp(a[0]);
  size_t bs = strtoull(a[1], 0, 10);
  double Fs = atof(a[2]), Fc = atof(a[3]), Ff = atof(a[4]), width = atof(a[5]);
  size_t order = strtoull(a[6], 0, 10), ss = strtoull(a[7], 0, 10);
  double oFs = atof(a[8]);
  int threads = atoi(a[9]); double min_s = atof(a[10]);
  size_t total = raw.size() / sizeof(C);
  size_t nbuf = total / bs;
  if (!nbuf) { fprintf(stderr, "input shorter than one buffer\n"); return 2; }

  struct Chain {
    Feed<C> feed; BB<S> bb; FMDemod<S, FMO> fm; Null<FMO> sink; Buffer<C> work;
    Chain(double Fc, double Ff, double w, size_t o, size_t ss, double oFs, double Fs, size_t bs)
      : bb(Fc, Ff, w, o, ss, oFs), work(bs) {
      feed.connect(&bb, true); bb.connect(&fm, true); fm.connect(&sink, true); feed.setup(Fs, bs);
    }
  };
  // every connect()/config() on the main thread (Logger/Queue singletons are lazily created)
  std::vector<Chain *> chains;
  for (int t = 0; t < threads; t++) chains.push_back(new Chain(Fc, Ff, width, order, ss, oFs, Fs, bs));

  std::vector<size_t> done(threads, 0);
  auto body = [&](int t, double secs) {
    Chain &c = *chains[t];
    auto t0 = std::chrono::steady_clock::now();
    size_t cnt = 0;
    for (;;) {
      for (size_t b = 0; b < nbuf; b++) {
        memcpy(c.work.data(), raw.data() + b * bs * sizeof(C), bs * sizeof(C));
        c.feed.push(c.work, true);
        cnt += bs;
      }
      double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (el >= secs) break;
    }
    done[t] = cnt;
  };
  body(0, 0.0);  // warm-up: one pass on chain 0
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++) th.emplace_back(body, t, min_s);
  for (auto &x : th) x.join();
  double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  size_t all = 0; for (int t = 0; t < threads; t++) all += done[t];
  printf("{\"msamples_per_s\": %.4f, \"threads\": %d, \"seconds\": %.3f, \"samples\": %zu, \"order\": %zu, \"sub_sample\": %zu}\n",
         all / el / 1e6, threads, el, all, chains[0]->bb.order(), chains[0]->bb.subSample());
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: ref_harness bb|ola|time ...\n"); return 2; }
  std::string cmd = argv[1];
  try {
    if (cmd == "bb" && argc == 14) {
      if (!strcmp(argv[2], "s16")) return run_bb<int16_t, int16_t>(argv + 3);
      if (!strcmp(argv[2], "s8")) return run_bb<int8_t, int16_t>(argv + 3);
    } else if (cmd == "ola" && argc == 8) {
      return run_ola(argv + 2);
    } else if (cmd == "rbb" && argc == 11) {
      return run_rbb<int16_t>(argv + 2);
    } else if (cmd == "rbb8" && argc == 11) {
      return run_rbb<int8_t>(argv + 2);
    } else if (cmd == "wav" && argc == 8) {
      std::string t = argv[2];
      if (t == "u8") return run_wav<uint8_t>(argv + 3);
      if (t == "s16") return run_wav<int16_t>(argv + 3);
      if (t == "cu8") return run_wav< std::complex<uint8_t> >(argv + 3);
      if (t == "cs16") return run_wav< std::complex<int16_t> >(argv + 3);
    } else if (cmd == "cast" && argc == 6) {
      if (!strcmp(argv[2], "cu8")) return run_cast< std::complex<uint8_t> >(argv + 3);
      if (!strcmp(argv[2], "cs8")) return run_cast< std::complex<int8_t> >(argv + 3);
    } else if (cmd == "castx" && argc == 7) {
      return run_castx(argv + 2);
    } else if (cmd == "deemph" && argc == 6) {
      return run_deemph(argv + 2);
    } else if (cmd == "time" && argc == 14) {
      if (!strcmp(argv[2], "s16")) return run_time<int16_t, int16_t>(argv + 3);
      if (!strcmp(argv[2], "s8")) return run_time<int8_t, int16_t>(argv + 3);
    }
  } catch (std::exception &e) {
    fprintf(stderr, "reference threw: %s\n", e.what());
    return 3;
  }
  fprintf(stderr, "bad arguments (argc=%d)\n", argc);
  return 2;
}
