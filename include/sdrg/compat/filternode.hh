// compatibility forwarder: lets sources written against libsdr (#include "filternode.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../filternode.hh"
