// compatibility forwarder: lets sources written against libsdr (#include "autocast.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../autocast.hh"
