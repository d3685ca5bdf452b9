// compatibility forwarder: lets sources written against libsdr (#include "sdr.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../sdr.hh"
