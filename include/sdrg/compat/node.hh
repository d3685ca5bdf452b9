// compatibility forwarder: lets sources written against libsdr (#include "node.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../node.hh"
