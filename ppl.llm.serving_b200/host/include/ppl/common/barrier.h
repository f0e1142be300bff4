// ppl::common::Barrier (EXTERNAL): Reset(n) then n threads Wait() (resource_manager.cc:344,416-417).
#ifndef B2LLM_SHIM_PPL_COMMON_BARRIER_H_
#define B2LLM_SHIM_PPL_COMMON_BARRIER_H_

#include <condition_variable>
#include <mutex>
#include <stdint.h>

namespace ppl { namespace common {

class Barrier final {
public:
    void Reset(uint32_t max_count) {
        std::lock_guard<std::mutex> g(mu_);
        max_ = max_count;
        arrived_ = 0;
        ++generation_;
    }
    void Wait() {
        std::unique_lock<std::mutex> lk(mu_);
        const uint64_t gen = generation_;
        if (++arrived_ >= max_) {
            arrived_ = 0;
            ++generation_;
            cv_.notify_all();
            return;
        }
        cv_.wait(lk, [&] { return generation_ != gen; });
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    uint32_t max_ = 0, arrived_ = 0;
    uint64_t generation_ = 0;
};

}} // namespace ppl::common

#endif
