// compatibility forwarder: lets sources written against libsdr (#include "exception.hh") build unchanged
// with -Iinclude/sdrg/compat
#include "../exception.hh"
