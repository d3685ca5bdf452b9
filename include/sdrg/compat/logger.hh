// compatibility forwarder: lets sources written against libsdr (#include "logger.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../logger.hh"
