// ppl::common::StaticThreadPool (EXTERNAL): N persistent worker threads, each always handling the same
// index (SURVEY.md 8b "Threading": the reference relies on thread i == TP rank i so that the CUDA device
// selected once by InitCudaEnv(i) stays current).  Calls used by the reference:
//   Init(n); GetNumThreads(); Run(f) = run f(nthr, ithr) on every worker and block until all returned
//   (utils.h:42-44); RunAsync(f) + Wait() (llm_generator.cc:622,696,738,784).
#ifndef B2LLM_SHIM_PPL_COMMON_THREADPOOL_H_
#define B2LLM_SHIM_PPL_COMMON_THREADPOOL_H_

#include "retcode.h"

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace ppl { namespace common {

class ThreadTask {
public:
    virtual ~ThreadTask() {}
};

class StaticThreadPool final {
public:
    StaticThreadPool() {}
    ~StaticThreadPool();

    RetCode Init(uint32_t thread_num);
    uint32_t GetNumThreads() const {
        return (uint32_t)workers_.size();
    }
    /** runs f(nthr, ithr) on every worker; returns when all are done */
    void Run(const std::function<void(uint32_t, uint32_t)>& f);
    /** same, without waiting; a previous round is waited for first */
    void RunAsync(const std::function<void(uint32_t, uint32_t)>& f);
    void Wait();

private:
    void WorkerLoop(uint32_t ithr);

    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_start_, cv_done_;
    std::function<void(uint32_t, uint32_t)> task_;
    uint64_t round_ = 0;      // incremented per submitted task
    uint32_t pending_ = 0;    // workers still running the current round
    bool stop_ = false;

    StaticThreadPool(const StaticThreadPool&) = delete;
    void operator=(const StaticThreadPool&) = delete;
};

}} // namespace ppl::common

#endif
