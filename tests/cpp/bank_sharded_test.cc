// A C/C++ application using more than one GPU through include/sdrg.h alone: the 2048-channel bank of
// BASELINE config 5 sharded over every visible device (sdrg_bank_sharded_*), host-pointer and
// device-pointer entry points, spot channels checked bit for bit against the oracle
// (oracle/sdr_oracle.c: IQBaseBand<int16_t> -> FMDemod / AMDemod out of place).  Needs a GPU; with one
// GPU the shards share it.
#include "sdrg.h"
#include "../../oracle/sdr_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)
#define OK(call) do { int rc_ = (call); if (rc_) { std::printf("FAIL %s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #call, rc_, sdrg_last_error()); return 1; } } while (0)

int main() {
  const size_t C = 2048, bs = 1 << 16, nbuf = 3;
  const double Fs = 100e6, width = 25e3, oFs = 48000.0;
  const size_t order = 15;
  int ndev = 0;
  OK(sdrg_device_count(&ndev));
  std::vector<int> devs;
  for (int d = 0; d < ndev; d++) devs.push_back(d);
  if (devs.size() < 2) devs = {0, 0, 0};

  std::vector<double> Fc(C);
  for (size_t k = 0; k < C; k++) Fc[k] = (double(k) - double(C / 2)) * (Fs / C);
  // input: a few carriers (amplitude 12 each) + noise +-8, then a full-scale noise tail
  std::vector<int16_t> x(2 * bs * nbuf);
  const size_t spot[8] = {0, 1, 1023, 1024, 1025, 2047, 389, 1707};
  uint32_t lcg = 4242;
  for (size_t n = 0; n < bs * nbuf; n++) {
    double re = 0, im = 0;
    for (size_t s = 0; s < 8; s++) { const double ph = 2 * M_PI * Fc[spot[s]] * (double(n) / Fs) + s; re += 12 * cos(ph); im += 12 * sin(ph); }
    lcg = lcg * 1664525u + 1013904223u; const int nr = int((lcg >> 16) % 17) - 8;
    lcg = lcg * 1664525u + 1013904223u; const int ni = int((lcg >> 16) % 17) - 8;
    if (n >= 2 * bs) { x[2 * n] = int16_t(lcg >> 16); lcg = lcg * 1664525u + 1013904223u; x[2 * n + 1] = int16_t(lcg >> 16); }
    else { x[2 * n] = int16_t(re) + nr; x[2 * n + 1] = int16_t(im) + ni; }
  }

  sdrg_bank_sharded *h = nullptr;
  OK(sdrg_bank_sharded_create(SDRG_T_S16, C, Fc.data(), nullptr, width, order, 1, oFs, devs.data(), devs.size(), &h));
  sdrg_config src = {SDRG_T_CS16, Fs, bs, 1}, out;
  OK(sdrg_bank_sharded_configure(h, &src, &out));
  CHECK(out.type == SDRG_T_CS16 && out.sample_rate == 48007.0);
  size_t nsh = 0, chans = 0;
  OK(sdrg_bank_sharded_info(h, &chans, &nsh, 0, nullptr, nullptr, nullptr, nullptr));
  CHECK(chans == C && nsh == devs.size());

  // (1) host pointers, two calls (2 buffers, then 1): the carried state lives in the shards
  size_t n1 = 0, n2 = 0, cap = 0;
  OK(sdrg_bank_sharded_outputs_for(h, bs * nbuf, &cap));
  const size_t stride = cap + 1;
  std::vector<int16_t> fm(C * stride, 0), am(C * stride, 0), fm2(C * stride, 0), am2(C * stride, 0);
  OK(sdrg_bank_sharded_process(h, x.data(), bs, 2, nullptr, fm.data(), am.data(), nullptr, stride, &n1));
  OK(sdrg_bank_sharded_process(h, x.data() + 2 * 2 * bs, bs, 1, nullptr, fm2.data(), am2.data(), nullptr, stride, &n2));
  CHECK(n1 + n2 == cap);

  // (2) device pointers on devices[0] after a fresh config()
  OK(sdrg_bank_sharded_configure(h, &src, &out));
  OK(sdrg_set_device(devs[0]));
  void *st = nullptr; OK(sdrg_stream_default(&st));
  void *d_all = nullptr;
  const size_t in_bytes = x.size() * 2, row_bytes = C * stride * 2;
  OK(sdrg_scratch(in_bytes + 2 * row_bytes + 64, &d_all));
  char *d_in = (char *)d_all, *d_fm = d_in + ((in_bytes + 15) & ~size_t(15)), *d_am = d_fm + row_bytes;
  std::vector<int16_t> zeros(C * stride, 0), dfm(C * stride), dam(C * stride);
  OK(sdrg_memcpy_h2d_async(d_in, x.data(), in_bytes, st));
  OK(sdrg_memcpy_h2d_async(d_fm, zeros.data(), row_bytes, st));
  size_t m = 0;
  OK(sdrg_bank_sharded_process_dev(h, d_in, bs, nbuf, nullptr, d_fm, d_am, nullptr, stride, &m, st));
  OK(sdrg_memcpy_d2h_async(dfm.data(), d_fm, row_bytes, st));
  OK(sdrg_memcpy_d2h_async(dam.data(), d_am, row_bytes, st));
  OK(sdrg_stream_synchronize(st));
  CHECK(m == cap);

  // oracle, per spot channel
  for (size_t s = 0; s < 8; s++) {
    const size_t c = spot[s];
    orc_iqbb *o = new orc_iqbb; int16_t last = 0;
    orc_iqbb_init(o, ORC_S16, Fc[c], Fc[c], width, order, 1, oFs);
    orc_iqbb_config(o, Fs, bs);
    std::vector<int16_t> bb(2 * (bs + 2)), ofm, oam; std::vector<char> skip;
    for (size_t b = 0; b < nbuf; b++) {
      const size_t n = orc_iqbb_process(o, &x[2 * b * bs], bs, bb.data());
      std::vector<int16_t> f(n + 1, 0), a(n + 1, 0);
      orc_fmdemod_s16(bb.data(), n, f.data(), &last);
      orc_amdemod_s16(bb.data(), n, a.data());
      for (size_t i = 0; i < n; i++) { ofm.push_back(f[i]); oam.push_back(a[i]); skip.push_back(i == 0); }
    }
    CHECK(ofm.size() == cap);
    size_t bad = 0;
    for (size_t i = 0; i < cap; i++) {
      const int16_t hf = i < n1 ? fm[c * stride + i] : fm2[c * stride + i - n1];
      const int16_t ha = i < n1 ? am[c * stride + i] : am2[c * stride + i - n1];
      if (ha != oam[i] || dam[c * stride + i] != oam[i]) bad++;
      if (skip[i]) { if (hf != 0 || dfm[c * stride + i] != 0) bad++; }     // out of place: element 0 of a buffer is never written
      else if (hf != ofm[i] || dfm[c * stride + i] != ofm[i]) bad++;
    }
    if (bad) std::printf("channel %zu: %zu mismatches\n", c, bad);
    CHECK(bad == 0);
    delete o;
  }
  OK(sdrg_bank_sharded_destroy(h));
  if (failures) { std::printf("bank_sharded_test: %d failure(s)\n", failures); return 1; }
  std::printf("bank_sharded_test: ok (%zu shards, %zu outputs per channel)\n", nsh, cap);
  return 0;
}
