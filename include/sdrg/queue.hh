// sdrg/queue.hh -- the message loop of the host-side mirror.
// Interface mirrored from src/queue.hh:53-216 / src/queue.cc: singleton Queue::get(); send() takes a
// reference on the buffer until the sink has handled it; start()/stop()/wait(); start, stop and idle
// delegates.  One loop thread delivers every queued buffer, so GPU nodes behind queued links are
// driven from that thread (the C ABI may be called from any host thread).
// Differences: std::thread/condition_variable instead of pthreads, and the queue length is only
// read under the lock (the reference reads it unlocked, src/queue.cc:95-97).
#ifndef SDRG_QUEUE_HH
#define SDRG_QUEUE_HH

#include <atomic>
#include <condition_variable>
#include <deque>
#include <list>
#include <mutex>
#include <thread>

#include "buffer.hh"
#include "logger.hh"

namespace sdr {

class SinkBase;

class DelegateInterface {
public:
  virtual ~DelegateInterface() {}
  virtual void operator()() = 0;
  virtual void *instance() = 0;
};

template <class T>
class Delegate : public DelegateInterface {
public:
  Delegate(T *instance, void (T::*func)(void)) : _instance(instance), _function(func) {}
  virtual ~Delegate() {}
  virtual void operator()() { (_instance->*_function)(); }
  virtual void *instance() { return _instance; }
protected:
  T *_instance;
  void (T::*_function)(void);
};

class Queue {
public:
  class Message {
  public:
    Message(const RawBuffer &buffer, SinkBase *sink, bool allow_overwrite)
      : _buffer(buffer), _sink(sink), _allow_overwrite(allow_overwrite) {}
    inline const RawBuffer &buffer() const { return _buffer; }
    inline RawBuffer &buffer() { return _buffer; }
    inline SinkBase *sink() const { return _sink; }
    inline bool allowOverwrite() const { return _allow_overwrite; }
  protected:
    RawBuffer _buffer;
    SinkBase *_sink;
    bool _allow_overwrite;
  };

  static Queue &get() { static Queue instance; return instance; }
  virtual ~Queue() { if (_thread.joinable()) { stop(); _thread.join(); } drain(); clear(_idle); clear(_onStart); clear(_onStop); }

  void send(const RawBuffer &buffer, SinkBase *sink, bool allow_overwrite = false) {
    std::lock_guard<std::mutex> lk(_lock);
    buffer.ref();
    _queue.push_back(Message(buffer, sink, allow_overwrite));
    _cond.notify_one();
  }
  void start() {
    if (_running.load()) return;
    if (_thread.joinable()) _thread.join();
    _running.store(true);
    // the worker inherits the starting thread's device: nodes created after sdrg_set_device(k) launch on
    // device k's stream when driven from the Queue thread as well
    int device = 0;
    sdrg_get_device(&device);
    _thread = std::thread(&Queue::threadMain, this, device);
  }
  void stop() { { std::lock_guard<std::mutex> lk(_lock); _running.store(false); } _cond.notify_all(); }
  void wait() { if (_thread.joinable()) _thread.join(); drain(); }
  bool isStopped() const { return !_running.load(); }
  bool isRunning() const { return _running.load(); }

  template <class T> void addIdle(T *instance, void (T::*function)(void)) { _idle.push_back(new Delegate<T>(instance, function)); }
  template <class T> void remIdle(T *instance) { remove(_idle, instance); }
  template <class T> void addStart(T *instance, void (T::*function)(void)) { _onStart.push_back(new Delegate<T>(instance, function)); }
  template <class T> void remStart(T *instance) { remove(_onStart, instance); }
  template <class T> void addStop(T *instance, void (T::*function)(void)) { _onStop.push_back(new Delegate<T>(instance, function)); }
  template <class T> void remStop(T *instance) { remove(_onStop, instance); }

protected:
  Queue() : _running(false) {}
  inline void deliver(Message &msg);   // defined in node.hh (needs SinkBase)
  void threadMain(int device) {
    sdrg_set_device(device);      // fails without a GPU: host-only graphs still run
    try { loop(); }
    catch (std::exception &err) {
      LogMessage msg(LOG_ERROR); msg << "Caught exception in thread: " << err.what() << " -> Stop thread.";
      Logger::get().log(msg);
    } catch (...) {
      Logger::get().log(LogMessage(LOG_ERROR, "Caught unknown exception in thread -> Stop thread."));
    }
    _running.store(false);
  }
  void loop() {
    Logger::get().log(LogMessage(LOG_DEBUG, "Queue started."));
    fire(_onStart);
    for (;;) {
      std::unique_lock<std::mutex> lk(_lock);
      if (_queue.empty()) {
        if (!_running.load()) break;
        lk.unlock();
        fire(_idle);
        lk.lock();
        _cond.wait(lk, [this] { return !_queue.empty() || !_running.load(); });
        if (_queue.empty()) { if (!_running.load()) break; continue; }
      }
      Message msg(_queue.front()); _queue.pop_front();
      lk.unlock();
      deliver(msg);
      msg.buffer().unref();
    }
    fire(_onStop);
    Logger::get().log(LogMessage(LOG_DEBUG, "Queue stopped."));
  }
  void drain() {
    std::lock_guard<std::mutex> lk(_lock);
    for (std::deque<Message>::iterator it = _queue.begin(); it != _queue.end(); ++it) it->buffer().unref();
    _queue.clear();
  }
  static void fire(std::list<DelegateInterface *> &l) {
    for (std::list<DelegateInterface *>::iterator it = l.begin(); it != l.end(); ++it) (**it)();
  }
  static void clear(std::list<DelegateInterface *> &l) {
    for (std::list<DelegateInterface *>::iterator it = l.begin(); it != l.end(); ++it) delete *it;
    l.clear();
  }
  template <class T> static void remove(std::list<DelegateInterface *> &l, T *instance) {
    for (std::list<DelegateInterface *>::iterator it = l.begin(); it != l.end();) {
      if ((*it)->instance() == (void *)instance) { delete *it; it = l.erase(it); } else ++it;
    }
  }

  std::atomic<bool> _running;
  std::thread _thread;
  std::mutex _lock;
  std::condition_variable _cond;
  std::deque<Message> _queue;
  std::list<DelegateInterface *> _idle, _onStart, _onStop;
};

}  // namespace sdr
#endif
