// compatibility forwarder: lets sources written against libsdr (#include "queue.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../queue.hh"
