// sdrg/wavfile.hh -- WavSink<Scalar> and WavSource: the file nodes either side of the receive chain
// (SURVEY.md 8f rank 4).  Host-only I/O with the interface and on-disk bytes of src/wavfile.hh:16-128
// and src/wavfile.cc:9-246: a 44-byte RIFF/WAVE PCM header that close() fills in, raw interleaved
// little-endian samples behind it.  The source's read buffer comes from RawBuffer's default policy
// (pinned + device mirrored from 4 KiB up), so a GPU node connected to it uploads straight from the
// page-locked read buffer; a GPU node's device-resident output reaching WavSink is copied back by
// Source::send first (WavSink is a host-only sink).
#ifndef SDRG_WAVFILE_HH
#define SDRG_WAVFILE_HH

#include "node.hh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>

namespace sdr {

namespace detail {
/** PCM layout of a Config type (channels, bits), 0 channels when WAV cannot hold it. */
inline void wavLayout(Config::Type t, uint16_t &channels, uint16_t &bits) {
  channels = 0; bits = 0;
  switch (t) {
    case Config::Type_u8: case Config::Type_s8: channels = 1; bits = 8; break;
    case Config::Type_cu8: case Config::Type_cs8: channels = 2; bits = 8; break;
    case Config::Type_u16: case Config::Type_s16: channels = 1; bits = 16; break;
    case Config::Type_cu16: case Config::Type_cs16: channels = 2; bits = 16; break;
    default: break;
  }
}
inline void put16(unsigned char *p, uint16_t v) { p[0] = (unsigned char)(v & 0xff); p[1] = (unsigned char)(v >> 8); }
inline void put32(unsigned char *p, uint32_t v) { put16(p, (uint16_t)(v & 0xffff)); put16(p + 2, (uint16_t)(v >> 16)); }
inline uint16_t get16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t get32(const unsigned char *p) { return (uint32_t)get16(p) | ((uint32_t)get16(p + 2) << 16); }
}  // namespace detail


/** Stores the received buffers as a PCM WAV file (src/wavfile.hh:16-128). */
template <class Scalar>
class WavSink : public Sink<Scalar> {
public:
  /** @throws ConfigError if the file cannot be opened or Scalar is not an 8/16-bit integer type. */
  WavSink(const std::string &filename) : Sink<Scalar>(), _fp(0), _frames(0), _rate(0), _channels(0), _bits(0) {
    _fp = std::fopen(filename.c_str(), "wb");
    if (!_fp) { ConfigError err; err << "Can not open wav file for output: " << filename; throw err; }
    const unsigned char zeros[kHeader] = {0};         // the header proper is written by close()
    std::fwrite(zeros, 1, kHeader, _fp);
    detail::wavLayout(Config::typeId<Scalar>(), _channels, _bits);
    if (0 == _channels) {
      std::fclose(_fp); _fp = 0;
      ConfigError err; err << "WAV format only allows (real) integer typed data."; throw err;
    }
  }
  virtual ~WavSink() { close(); }

  virtual void config(const Config &src_cfg) {
    if (!src_cfg.hasType() || !src_cfg.hasSampleRate()) return;
    if (Config::typeId<Scalar>() != src_cfg.type()) {
      ConfigError err;
      err << "Can not configure WavSink: Invalid buffer type " << src_cfg.type() << ", expected " << Config::typeId<Scalar>();
      throw err;
    }
    _rate = (uint32_t)src_cfg.sampleRate();
  }

  /** Completes the header and closes the file.  Field for field what the reference writes,
   * including its RIFF chunk size of 36 + 2*frames whatever the frame size (wavfile.hh:91). */
  void close() {
    if (!_fp) return;
    const uint16_t frame_bytes = (uint16_t)(_channels * (_bits / 8));
    unsigned char h[kHeader];
    std::memcpy(h, "RIFF", 4);      detail::put32(h + 4, 36u + 2u * _frames);
    std::memcpy(h + 8, "WAVE", 4);
    std::memcpy(h + 12, "fmt ", 4); detail::put32(h + 16, 16);
    detail::put16(h + 20, 1);       detail::put16(h + 22, _channels);
    detail::put32(h + 24, _rate);   detail::put32(h + 28, (uint32_t)frame_bytes * _rate);
    detail::put16(h + 32, frame_bytes); detail::put16(h + 34, _bits);
    std::memcpy(h + 36, "data", 4); detail::put32(h + 40, (uint32_t)frame_bytes * _frames);
    std::fseek(_fp, 0, SEEK_SET);
    std::fwrite(h, 1, kHeader, _fp);
    std::fclose(_fp); _fp = 0;
  }

  virtual void process(const Buffer<Scalar> &buffer, bool /*allow_overwrite*/) {
    if (!_fp) return;
    std::fwrite(buffer.data(), sizeof(Scalar), buffer.size(), _fp);
    _frames += (uint32_t)buffer.size();
  }

protected:
  static const size_t kHeader = 44;
  FILE *_fp;
  uint32_t _frames, _rate;
  uint16_t _channels, _bits;
};


/** Reads 8- or 16-bit, 1- or 2-channel PCM WAV files (src/wavfile.cc:9-246): u8, s16, cu8, cs16. */
class WavSource : public Source {
public:
  WavSource(size_t buffer_size = 1024)
    : Source(), _fp(0), _buffer_size(buffer_size), _frame_count(0), _frame_bytes(0), _type(Config::Type_UNDEFINED),
      _sample_rate(0), _frames_left(0) {}
  WavSource(const std::string &filename, size_t buffer_size = 1024)
    : Source(), _fp(0), _buffer_size(buffer_size), _frame_count(0), _frame_bytes(0), _type(Config::Type_UNDEFINED),
      _sample_rate(0), _frames_left(0) { open(filename); }
  virtual ~WavSource() { closeFile(); _buffer.unref(); }

  bool isOpen() const { return 0 != _fp; }
  bool isReal() const { return (Config::Type_u8 == _type) || (Config::Type_s16 == _type); }
  size_t frameCount() const { return _frame_count; }

  /** Parses the header and publishes the source's config; an unopenable file is not an error
   * (isOpen() stays false, wavfile.cc:37), a malformed one throws RuntimeError. */
  void open(const std::string &filename) {
    closeFile();
    _fp = std::fopen(filename.c_str(), "rb");
    if (!_fp) return;
    unsigned char h[12];
    if (!readExact(h, 12) || std::memcmp(h, "RIFF", 4) || std::memcmp(h + 8, "WAVE", 4)) {
      closeFile(); RuntimeError err; err << "File '" << filename << "' is not a WAV file."; throw err;
    }
    unsigned char ck[8];
    if (!readExact(ck, 8) || std::memcmp(ck, "fmt ", 4)) {
      closeFile(); RuntimeError err; err << "'File 'fmt' header missing in file " << filename << "' @" << 16; throw err;
    }
    const uint32_t fmt_size = detail::get32(ck + 4);
    unsigned char f[16];
    if (!readExact(f, 16)) { closeFile(); RuntimeError err; err << "File '" << filename << "' is not a WAV file."; throw err; }
    const uint16_t format = detail::get16(f), channels = detail::get16(f + 2);
    const uint32_t rate = detail::get32(f + 4);
    const uint16_t align = detail::get16(f + 12), bits = detail::get16(f + 14);
    if (1 != format) {
      closeFile(); RuntimeError err;
      err << "Unsupported WAV data format: " << format << " of file " << filename << ". Expected " << 1; throw err;
    }
    if ((1 != channels) && (2 != channels)) {
      closeFile(); RuntimeError err;
      err << "Unsupported number of chanels: " << channels << " of file " << filename << ". Expected 1 or 2."; throw err;
    }
    if ((16 != bits) && (8 != bits)) {
      closeFile(); RuntimeError err;
      err << "Unsupported sample format: " << bits << "b of file " << filename << ". Expected 16b or 8b."; throw err;
    }
    if (align != channels * (bits / 8)) {
      closeFile(); RuntimeError err;
      err << "Unsupported alignment: " << align << "byte of file " << filename << ". Expected " << (bits / 8) << "byte."; throw err;
    }
    // chunks after "fmt " are skipped until "data" (the reference spins forever on a file without one)
    long offset = 12 + 8 + (long)fmt_size;
    uint32_t data_bytes = 0;
    for (;;) {
      if (std::fseek(_fp, offset, SEEK_SET) || !readExact(ck, 8)) {
        closeFile(); RuntimeError err; err << "WAV file '" << filename << "' contains no 'data' chunk."; throw err;
      }
      if (0 == std::memcmp(ck, "data", 4)) { data_bytes = detail::get32(ck + 4); break; }
      offset += 8 + (long)detail::get32(ck + 4);
    }
    _frame_bytes = (size_t)channels * (bits / 8);
    _frame_count = data_bytes / _frame_bytes;
    if (1 == channels) _type = (8 == bits) ? Config::Type_u8 : Config::Type_s16;
    else _type = (8 == bits) ? Config::Type_cu8 : Config::Type_cs16;
    _sample_rate = rate;
    _frames_left = _frame_count;

    LogMessage msg(LOG_DEBUG);
    msg << "Configured WavSource:" << std::endl << " file: " << filename << std::endl << " type: " << _type << std::endl
        << " sample-rate: " << _sample_rate << std::endl << " frame-count: " << _frame_count << std::endl
        << " duration: " << _frame_count / _sample_rate << "s" << std::endl << " buffer-size: " << _buffer_size;
    Logger::get().log(msg);

    if (!_buffer.isEmpty()) _buffer.unref();
    _buffer = RawBuffer(_buffer_size * _frame_bytes);
    this->setConfig(Config(_type, _sample_rate, _buffer_size, 1));
  }

  void close() { closeFile(); _frames_left = 0; }

  /** Reads and sends the next buffer (at most buffer_size frames); at the end of the data the file
   * is closed and end-of-stream signalled (wavfile.cc:203-210). */
  void next() {
    if (0 == _frames_left) { closeFile(); signalEOS(); return; }
    const size_t n_frames = std::min(_frames_left, _buffer_size);
    const size_t got = _fp ? std::fread(_buffer.ptr(), _frame_bytes, n_frames, _fp) : 0;
    if (got < n_frames) std::memset(_buffer.ptr() + got * _frame_bytes, 0, (n_frames - got) * _frame_bytes);   // truncated file
    _frames_left -= n_frames;
    this->send(RawBuffer(_buffer, 0, n_frames * _frame_bytes), true);
  }

protected:
  bool readExact(void *p, size_t n) { return _fp && std::fread(p, 1, n, _fp) == n; }
  void closeFile() { if (_fp) { std::fclose(_fp); _fp = 0; } }

  FILE *_fp;
  RawBuffer _buffer;
  size_t _buffer_size, _frame_count, _frame_bytes;
  Config::Type _type;
  double _sample_rate;
  size_t _frames_left;
};

}  // namespace sdr
#endif
