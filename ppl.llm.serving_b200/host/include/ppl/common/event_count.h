// ppl::common::EventCount (EXTERNAL): the two-phase wait the generator thread sleeps on
// (llm_generator.cc:342-366: PrepareWait -> re-check -> CancelWait | CommitWait; producers NotifyOne).
// A ticket taken by PrepareWait is stale as soon as any Notify* happened after it, so a notification that
// races with the re-check is never lost.
#ifndef B2LLM_SHIM_PPL_COMMON_EVENT_COUNT_H_
#define B2LLM_SHIM_PPL_COMMON_EVENT_COUNT_H_

#include <condition_variable>
#include <mutex>
#include <stdint.h>

namespace ppl { namespace common {

class EventCount final {
public:
    typedef uint64_t Key;

    Key PrepareWait() {
        std::lock_guard<std::mutex> g(mu_);
        ++waiters_;
        return epoch_;
    }
    void CancelWait() {
        std::lock_guard<std::mutex> g(mu_);
        --waiters_;
    }
    void CommitWait(Key key) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != key; });
        --waiters_;
    }
    void NotifyOne() {
        std::lock_guard<std::mutex> g(mu_);
        ++epoch_;
        cv_.notify_all(); // every committed waiter holds a key older than the new epoch
    }
    void NotifyAll() {
        NotifyOne();
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    uint64_t epoch_ = 0;
    uint32_t waiters_ = 0;
};

}} // namespace ppl::common

#endif
