// sdrg/logger.hh -- minimal logger with the reference's surface (src/logger.hh:12-111):
// LogLevel, LogMessage (a stringstream with a level), LogHandler, StreamLogHandler, Logger::get().
// The hot path logs only from config().
#ifndef SDRG_LOGGER_HH
#define SDRG_LOGGER_HH

#include <iostream>
#include <list>
#include <mutex>
#include <sstream>
#include <string>

namespace sdr {

typedef enum { LOG_DEBUG = 0, LOG_INFO, LOG_WARNING, LOG_ERROR } LogLevel;

class LogMessage : public std::stringstream {
public:
  LogMessage(LogLevel level, const std::string &msg = "") : _level(level) { (*this) << msg; }
  LogMessage(const LogMessage &o) : std::stringstream(), _level(o._level) { (*this) << o.str(); }
  virtual ~LogMessage() {}
  LogLevel level() const { return _level; }
  std::string message() const { return this->str(); }
protected:
  LogLevel _level;
};

class LogHandler {
public:
  virtual ~LogHandler() {}
  virtual void handle(const LogMessage &msg) = 0;
};

class StreamLogHandler : public LogHandler {
public:
  StreamLogHandler(std::ostream &stream, LogLevel level) : _stream(stream), _level(level) {}
  virtual void handle(const LogMessage &msg) {
    if (msg.level() < _level) return;
    static const char *names[] = {"DEBUG", "INFO", "WARN", "ERROR"};
    _stream << names[msg.level()] << ": " << msg.message() << std::endl;
  }
protected:
  std::ostream &_stream;
  LogLevel _level;
};

class Logger {
public:
  static Logger &get() { static Logger instance; return instance; }
  void log(const LogMessage &message) {
    std::lock_guard<std::mutex> lk(_mu);
    for (std::list<LogHandler *>::iterator it = _handlers.begin(); it != _handlers.end(); ++it) (*it)->handle(message);
  }
  void addHandler(LogHandler *handler) { std::lock_guard<std::mutex> lk(_mu); _handlers.push_back(handler); }  // takes ownership
  ~Logger() { for (std::list<LogHandler *>::iterator it = _handlers.begin(); it != _handlers.end(); ++it) delete *it; }
protected:
  Logger() {}
  std::mutex _mu;
  std::list<LogHandler *> _handlers;
};

}  // namespace sdr
#endif
