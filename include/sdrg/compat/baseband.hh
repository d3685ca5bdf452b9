// compatibility forwarder: lets sources written against libsdr (#include "baseband.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../baseband.hh"
