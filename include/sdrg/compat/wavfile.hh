// libsdr header name -> B200-native implementation (src/wavfile.hh)
#include "../wavfile.hh"
