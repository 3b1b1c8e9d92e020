// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for Taskflow v3.8.0
// (taskflow/taskflow @ d8c49c64, pinned by /root/reference/icicle/backend/cpu/CMakeLists.txt:19-24),
// which is an un-vendored dependency of the reference CPU backend. The reference uses it purely as a
// thread pool -- it carries no arithmetic -- through exactly this surface:
//   tf::Taskflow::emplace(callable), tf::Taskflow::clear()
//   tf::Executor(), tf::Executor(n), tf::Executor::run(taskflow).wait()
// (call sites: backend/cpu/src/curve/cpu_msm.hpp:175-176,228-229,246,321; backend/cpu/include/ntt_cpu.h:79-80,106,117;
//  backend/cpu/src/field/cpu_vec_ops.cpp:645-674; backend/cpu/src/field/cpu_matrix_ops.cpp:80-124).
// Written from scratch for this repo; not a copy of Taskflow.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace tf {

  class Task
  {
  public:
    Task() = default;
  };

  class Taskflow
  {
  public:
    template <typename F>
    Task emplace(F&& f)
    {
      m_tasks.emplace_back(std::forward<F>(f));
      return Task();
    }
    void clear() { m_tasks.clear(); }
    bool empty() const { return m_tasks.empty(); }
    size_t num_tasks() const { return m_tasks.size(); }
    std::vector<std::function<void()>> m_tasks;
  };

  class Executor
  {
  public:
    class RunHandle
    {
    public:
      explicit RunHandle(Executor* e) : m_exec(e) {}
      void wait() { m_exec->wait_all(); }
      void get() { m_exec->wait_all(); }

    private:
      Executor* m_exec;
    };

    explicit Executor(size_t n = std::thread::hardware_concurrency())
    {
      if (n == 0) n = 1;
      m_threads.reserve(n);
      for (size_t i = 0; i < n; ++i)
        m_threads.emplace_back([this]() { worker_loop(); });
    }
    ~Executor()
    {
      {
        std::lock_guard<std::mutex> lk(m_mu);
        m_stop = true;
      }
      m_cv.notify_all();
      for (auto& t : m_threads)
        t.join();
    }
    Executor(const Executor&) = delete;
    Executor& operator=(const Executor&) = delete;

    size_t num_workers() const { return m_threads.size(); }

    RunHandle run(Taskflow& tf)
    {
      {
        std::lock_guard<std::mutex> lk(m_mu);
        m_cur = &tf.m_tasks;
        m_next = 0;
        m_pending = tf.m_tasks.size();
      }
      m_cv.notify_all();
      return RunHandle(this);
    }

  private:
    void wait_all()
    {
      std::unique_lock<std::mutex> lk(m_mu);
      m_done_cv.wait(lk, [this]() { return m_pending == 0; });
      m_cur = nullptr;
    }
    void worker_loop()
    {
      std::unique_lock<std::mutex> lk(m_mu);
      for (;;) {
        m_cv.wait(lk, [this]() { return m_stop || (m_cur && m_next < m_cur->size()); });
        if (m_stop) return;
        while (m_cur && m_next < m_cur->size()) {
          std::function<void()>& fn = (*m_cur)[m_next++];
          lk.unlock();
          fn();
          lk.lock();
          if (--m_pending == 0) m_done_cv.notify_all();
        }
      }
    }

    std::vector<std::thread> m_threads;
    std::mutex m_mu;
    std::condition_variable m_cv, m_done_cv;
    std::vector<std::function<void()>>* m_cur = nullptr;
    size_t m_next = 0;
    size_t m_pending = 0;
    bool m_stop = false;
  };

} // namespace tf
