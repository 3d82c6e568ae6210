// util.h -- the MRPT infrastructure pieces the front-end uses, in miniature:
// mrpt::WorkerThreadsPool (FIFO, LidarOdometry.h:167-172) and
// mrpt::system::CTimeLogger `profiler_` with the reference's section names
// (SURVEY.md section 5).
#pragma once
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace mola
{
class WorkerThreadsPool
{
   public:
    explicit WorkerThreadsPool(size_t n = 1) { resize(n); }
    ~WorkerThreadsPool() { clear(); }
    void resize(size_t n)
    {
        clear();
        stop_ = false;
        for (size_t i = 0; i < n; i++) threads_.emplace_back([this] { loop(); });
    }
    void clear()
    {
        {
            std::unique_lock<std::mutex> lk(mtx_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_)
            if (t.joinable()) t.join();
        threads_.clear();
    }
    void enqueue(std::function<void()> f)
    {
        {
            std::unique_lock<std::mutex> lk(mtx_);
            tasks_.push_back(std::move(f));
        }
        cv_.notify_one();
    }
    size_t pendingTasks()
    {
        std::unique_lock<std::mutex> lk(mtx_);
        return tasks_.size();
    }
    /** blocks until the queue is empty and every worker is idle */
    void waitIdle()
    {
        std::unique_lock<std::mutex> lk(mtx_);
        idle_cv_.wait(lk, [this] { return tasks_.empty() && busy_ == 0; });
    }

   private:
    void loop()
    {
        for (;;)
        {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(mtx_);
                cv_.wait(lk, [this] { return stop_ || !tasks_.empty(); });
                if (stop_ && tasks_.empty()) return;
                f = std::move(tasks_.front());
                tasks_.pop_front();
                busy_++;
            }
            f();
            {
                std::unique_lock<std::mutex> lk(mtx_);
                busy_--;
            }
            idle_cv_.notify_all();
        }
    }
    std::vector<std::thread>          threads_;
    std::deque<std::function<void()>> tasks_;
    std::mutex                        mtx_;
    std::condition_variable           cv_, idle_cv_;
    bool                              stop_ = false;
    size_t                            busy_ = 0;
};

class TimeLogger
{
   public:
    struct Stat
    {
        size_t n = 0;
        double total = 0, min = 1e300, max = 0;
    };
    // Open sections are kept per (name, thread): run_one_icp / doCheckForNonAdjacentKFs run concurrently on the
    // pool threads.  A section left on another thread than it was entered on (delay_onNewObs_to_process:
    // entered by the data source's thread, left by the worker) closes the oldest open one of that name.
    void enter(const std::string& name)
    {
        std::lock_guard<std::mutex> lk(mtx_);
        open_.emplace(name, Open{std::this_thread::get_id(), now()});
    }
    double leave(const std::string& name)
    {
        std::lock_guard<std::mutex> lk(mtx_);
        auto                        range = open_.equal_range(name);
        if (range.first == range.second) return 0;
        auto pick = range.second;
        for (auto it = range.first; it != range.second; ++it)
            if (it->second.tid == std::this_thread::get_id()) pick = it;  // the innermost of this thread
        if (pick == range.second)
        {
            pick = range.first;
            for (auto it = range.first; it != range.second; ++it)
                if (it->second.t0 < pick->second.t0) pick = it;
        }
        const double dt = now() - pick->second.t0;
        open_.erase(pick);
        add(name, dt);
        return dt;
    }
    void registerUserMeasure(const std::string& name, double v)
    {
        std::lock_guard<std::mutex> lk(mtx_);
        add(name, v);
    }
    std::map<std::string, Stat> stats()
    {
        std::lock_guard<std::mutex> lk(mtx_);
        return stats_;
    }
    void clear()
    {
        std::lock_guard<std::mutex> lk(mtx_);
        stats_.clear();
        open_.clear();
    }
    static double now()
    {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }

   private:
    void add(const std::string& name, double v)
    {
        Stat& s = stats_[name];
        s.n++, s.total += v;
        if (v < s.min) s.min = v;
        if (v > s.max) s.max = v;
    }
    std::mutex                    mtx_;
    std::map<std::string, Stat>   stats_;
    struct Open
    {
        std::thread::id tid;
        double          t0;
    };
    std::multimap<std::string, Open> open_;
};

struct ProfilerEntry
{
    TimeLogger& tl;
    std::string name;
    bool        open = true;
    ProfilerEntry(TimeLogger& t, std::string n) : tl(t), name(std::move(n)) { tl.enter(name); }
    void stop()
    {
        if (open) tl.leave(name);
        open = false;
    }
    ~ProfilerEntry() { stop(); }
};

}  // namespace mola
