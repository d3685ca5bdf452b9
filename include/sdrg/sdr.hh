// sdrg/sdr.hh -- umbrella header of the B200-native libsdr receive chain (drop-in for the subset of
// src/sdr.hh that lies on the hot path: node runtime, buffers, IQBaseBand, demodulators, FFT filter).
#ifndef SDRG_SDR_HH
#define SDRG_SDR_HH
#include "exception.hh"
#include "logger.hh"
#include "traits.hh"
#include "buffer.hh"
#include "queue.hh"
#include "node.hh"
#include "baseband.hh"
#include "demod.hh"
#include "autocast.hh"
#include "fftplan.hh"
#include "filternode.hh"
#include "wavfile.hh"
#endif
