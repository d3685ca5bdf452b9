// compatibility forwarder: lets sources written against libsdr (#include "fftplan.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../fftplan.hh"
