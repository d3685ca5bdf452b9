// Boundary-conformance tests of the Buffer family: a restatement of the reference's own
// test/buffertest.cc:9-122 and test/coretest.cc:10-25 against include/sdrg/buffer.hh.
// Runs on the host only (no GPU needed: without a device the storage falls back to the heap).
#include "sdrg/sdr.hh"
#include <cstdio>
#include <list>

using namespace sdr;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

struct Pool : public BufferOwner {
  int unused_calls = 0;
  virtual void bufferUnused(const RawBuffer &) { ++unused_calls; }
};

int main() {
  {   // reference counting (buffertest.cc:9-51)
    Buffer<int8_t> a(3);
    CHECK(a.refCount() == 1); CHECK(a.isUnused());
    { Buffer<int8_t> b(a); CHECK(a.refCount() == 1); CHECK(b.refCount() == 1); CHECK(b.isUnused()); }
    { Buffer<int8_t> b(a); b.ref(); CHECK(a.refCount() == 2); CHECK(!a.isUnused()); CHECK(!b.isUnused()); b.unref(); }
    CHECK(a.refCount() == 1); CHECK(a.isUnused());
    std::list<RawBuffer> l; l.push_back(a);
    CHECK(a.refCount() == 1); CHECK(l.back().refCount() == 1); CHECK(l.back().isUnused());
    l.pop_back(); CHECK(a.refCount() == 1);
    a.unref(); CHECK(a.isEmpty());
  }
  {   // reinterpretation (buffertest.cc:54-70)
    Buffer<int8_t> r(4); r[0] = 1; r[1] = 2; r[2] = 3; r[3] = 4;
    Buffer< std::complex<int8_t> > c(r);
    CHECK(r.size() / 2 == c.size());
    CHECK(c[0] == std::complex<int8_t>(1, 2)); CHECK(c[1] == std::complex<int8_t>(3, 4));
    CHECK(c.head(1).size() == 1); CHECK(c.tail(1)[0] == std::complex<int8_t>(3, 4)); CHECK(c.sub(1, 2).isEmpty());
    r.unref();
  }
  {   // raw ring buffer (buffertest.cc:73-122)
    RawBuffer a(3), b(3); RawRingBuffer ring(3);
    memcpy(a.data(), "abc", 3);
    CHECK(ring.bytesLen() == 0); CHECK(ring.bytesFree() == 3);
    CHECK(ring.put(RawBuffer(a, 0, 1))); CHECK(ring.bytesLen() == 1); CHECK(ring.bytesFree() == 2);
    CHECK(ring.put(RawBuffer(a, 1, 2))); CHECK(ring.bytesLen() == 3); CHECK(ring.bytesFree() == 0);
    CHECK(!ring.put(a));
    CHECK(ring.take(b, 1)); CHECK(ring.bytesLen() == 2); CHECK(*(b.data()) == 'a');
    CHECK(ring.take(b, 1)); CHECK(ring.bytesLen() == 1); CHECK(*(b.data()) == 'b');
    CHECK(ring.put(RawBuffer(a, 0, 2))); CHECK(ring.bytesLen() == 3);
    CHECK(ring.take(b, 3)); CHECK(ring.bytesLen() == 0); CHECK(0 == memcmp(b.data(), "cab", 3));
    a.unref(); b.unref(); ring.unref();
  }
  {   // owner notification and the pool (buffer.hh:287-352; resize() makes its buffers available)
    Pool p; RawBuffer x(16, &p); x.ref(); CHECK(p.unused_calls == 0); x.unref(); CHECK(p.unused_calls == 1); x.unref();
    BufferSet<int16_t> set(0, 8); CHECK(!set.hasBuffer()); set.resize(2); CHECK(set.hasBuffer());
    Buffer<int16_t> b1 = set.getBuffer(), b2 = set.getBuffer(); CHECK(!set.hasBuffer());
    b1.ref(); b1.unref(); CHECK(set.hasBuffer()); (void)b2;
  }
  {   // shift semantics the fixed-point path relies on (coretest.cc:10-25)
    int a = 128, b = -128;
    CHECK((a >> 1) == 64); CHECK((a << 1) == 256); CHECK((b >> 1) == -64); CHECK((b << 1) == -256);
  }
  {   // config propagation and exceptions (node.cc:86-114)
    struct Probe : public Sink<int16_t> { int configs = 0; Config last;
      virtual void config(const Config &c) { ++configs; last = c; } virtual void process(const Buffer<int16_t> &, bool) {} } probe;
    Proxy px; px.connect(&probe, true); CHECK(probe.configs == 1);
    px.config(Config(Config::Type_s16, 48e3, 64, 1)); CHECK(probe.configs == 2); CHECK(probe.last.bufferSize() == 64);
    px.config(Config(Config::Type_s16, 48e3, 64, 1)); CHECK(probe.configs == 2);   // unchanged: not re-propagated
    bool threw = false;
    try { ConfigError e; e << "x" << 1; throw e; } catch (SDRError &e) { threw = std::string(e.what()) == "x1"; }
    CHECK(threw);
  }
  {   // BlockingSource (src/node.hh:267-311, node.cc:137-190): idle-driven and parallel, EOS stops the queue
    // The Queue fires the idle signal once each time it runs dry and then sleeps until a buffer is queued
    // (src/queue.cc:106-115), so an idle-driven source must feed a QUEUED link from next(), as RTL/Port sources do.
    struct Count : public Sink<int16_t> { int buffers = 0; virtual void config(const Config &) {} virtual void process(const Buffer<int16_t> &, bool) { ++buffers; } };
    struct Counter : public BlockingSource {
      int calls = 0, limit; Buffer<int16_t> buf;
      Counter(bool parallel, int lim) : BlockingSource(parallel, true, true), limit(lim), buf(8) {}
      virtual void next() {
        if (++calls >= limit) { _is_active = false; signalEOS(); return; }
        send(buf, false);
      }
    };
    {
      Counter idle(false, 5); Count sink; idle.connect(&sink, false);
      CHECK(!idle.isActive()); idle.start(); CHECK(idle.isActive());
      Queue::get().start(); Queue::get().wait();        // next() runs on the idle signal until EOS stops the queue
      CHECK(idle.calls == 5); CHECK(sink.buffers == 4); CHECK(!idle.isActive()); CHECK(Queue::get().isStopped());
    }
    {
      Counter par(true, 3); Count sink; par.connect(&sink, false);
      Queue::get().start();
      while (!Queue::get().isRunning()) {}
      par.start();                                      // its own thread calls next() while the queue runs
      Queue::get().wait();
      par.stop();
      CHECK(par.calls == 3); CHECK(sink.buffers == 2);
    }
  }
  std::printf(failures ? "buffer_test: %d FAILED\n" : "buffer_test: ok\n", failures);
  return failures ? 1 : 0;
}
