// sdrg/exception.hh -- error types of the host-side mirror.
// Interface mirrored: src/exception.hh:10-45 (SDRError : std::exception, std::stringstream;
// ConfigError, RuntimeError).  `err << "text"; throw err;` keeps working.
#ifndef SDRG_EXCEPTION_HH
#define SDRG_EXCEPTION_HH

#include <exception>
#include <sstream>
#include <string>

namespace sdr {

class SDRError : public std::exception, public std::stringstream {
public:
  SDRError() {}
  SDRError(const SDRError &o) : std::exception(), std::stringstream() { (*this) << o.str(); }
  virtual ~SDRError() throw() {}
  virtual const char *what() const throw() { _what = this->str(); return _what.c_str(); }
private:
  mutable std::string _what;
};

class ConfigError : public SDRError {
public:
  ConfigError() {}
  ConfigError(const ConfigError &o) : SDRError(o) {}
  virtual ~ConfigError() throw() {}
};

class RuntimeError : public SDRError {
public:
  RuntimeError() {}
  RuntimeError(const RuntimeError &o) : SDRError(o) {}
  virtual ~RuntimeError() throw() {}
};

}  // namespace sdr
#endif
