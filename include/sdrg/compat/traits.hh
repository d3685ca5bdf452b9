// compatibility forwarder: lets sources written against libsdr (#include "traits.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../traits.hh"
