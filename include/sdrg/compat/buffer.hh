// compatibility forwarder: lets sources written against libsdr (#include "buffer.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../buffer.hh"
