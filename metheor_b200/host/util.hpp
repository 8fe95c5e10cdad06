// util.hpp — small host utilities: thread pool, little-endian readers, error type, wall clock.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace mthh {

// An error that maps to the reference's process behaviour: Rust panics exit with status 101 and print the panic
// message on stderr (bamutil.rs:8, readutil.rs:46-50,356); clap usage errors exit with status 2.
struct HostError {
    int status;
    std::string msg;
};

// multi-GPU sharding (run.cpp): the sites [lo, hi) of contig tid are owned by one rank
struct ShardInterval {
    int32_t tid;
    int64_t lo, hi;
};
std::vector<std::vector<ShardInterval>> plan_bins(const std::vector<int64_t>& ref_len, int world, int64_t region_cost = 0);
int64_t shard_region_cost_env();  // METHEOR_SHARD_REGION_COST (bases), 0 if unset
std::vector<std::vector<ShardInterval>> plan_contigs(const std::vector<int64_t>& ref_len, int world);

inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline int32_t le32(const uint8_t* p) {
    uint32_t v;
    memcpy(&v, p, 4);  // x86-64 / aarch64 are little-endian, as is BAM
    return (int32_t)v;
}

// Fork-join pool: run(n, fn) calls fn(task, worker) for task in [0, n) on the pool's threads (dynamic scheduling)
// and returns when all are done.  The caller thread participates as worker 0.
class ThreadPool {
public:
    explicit ThreadPool(int n_threads) : n_(n_threads < 1 ? 1 : n_threads) {
        for (int w = 1; w < n_; w++) threads_.emplace_back([this, w] { loop(w); });
    }
    ~ThreadPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return n_; }

    void run(int64_t n_tasks, const std::function<void(int64_t, int)>& fn) {
        if (n_tasks <= 0) return;
        if (n_ == 1 || n_tasks == 1) {
            for (int64_t i = 0; i < n_tasks; i++) fn(i, 0);
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            n_tasks_ = n_tasks;
            next_.store(0);
            pending_ = n_ - 1;
            gen_++;
        }
        cv_.notify_all();
        work(0);
        std::unique_lock<std::mutex> g(m_);
        done_cv_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void work(int w) {
        for (;;) {
            int64_t i = next_.fetch_add(1);
            if (i >= n_tasks_) break;
            (*fn_)(i, w);
        }
    }
    void loop(int w) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work(w);
            {
                std::lock_guard<std::mutex> g(m_);
                pending_--;
            }
            done_cv_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int64_t, int)>* fn_ = nullptr;
    int64_t n_tasks_ = 0;
    std::atomic<int64_t> next_{0};
    int pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

}  // namespace mthh
