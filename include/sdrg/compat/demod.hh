// compatibility forwarder: lets sources written against libsdr (#include "demod.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../demod.hh"
