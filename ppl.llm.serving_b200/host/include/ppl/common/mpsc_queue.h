// ppl::common::MPSCQueue (EXTERNAL): intrusive multi-producer single-consumer queue
// (mpsc_request_scheduler.h:58-88: Push(node); Pop(&is_empty) may return nullptr with is_empty == false
// while a producer is between its two stores -- the caller spins).  Vyukov's intrusive MPSC algorithm.
#ifndef B2LLM_SHIM_PPL_COMMON_MPSC_QUEUE_H_
#define B2LLM_SHIM_PPL_COMMON_MPSC_QUEUE_H_

#include <atomic>

namespace ppl { namespace common {

class MPSCQueue final {
public:
    struct Node {
        std::atomic<Node*> mpsc_next_{nullptr};
        virtual ~Node() {}
    };

    MPSCQueue() : head_(&stub_), tail_(&stub_) {}

    void Push(Node* n) {
        n->mpsc_next_.store(nullptr, std::memory_order_relaxed);
        Node* prev = head_.exchange(n, std::memory_order_acq_rel);
        prev->mpsc_next_.store(n, std::memory_order_release);
    }

    /** consumer side only */
    Node* Pop(bool* is_empty) {
        Node* tail = tail_;
        Node* next = tail->mpsc_next_.load(std::memory_order_acquire);
        if (tail == &stub_) {
            if (!next) {
                *is_empty = (head_.load(std::memory_order_acquire) == &stub_);
                return nullptr;
            }
            tail_ = next;
            tail = next;
            next = next->mpsc_next_.load(std::memory_order_acquire);
        }
        if (next) {
            tail_ = next;
            *is_empty = false;
            return tail;
        }
        if (tail != head_.load(std::memory_order_acquire)) {
            *is_empty = false; // a producer is mid-push
            return nullptr;
        }
        Push(&stub_);
        next = tail->mpsc_next_.load(std::memory_order_acquire);
        if (next) {
            tail_ = next;
            *is_empty = false;
            return tail;
        }
        *is_empty = false;
        return nullptr;
    }

private:
    Node stub_;
    std::atomic<Node*> head_; // producers exchange here
    Node* tail_;              // consumer reads here

    MPSCQueue(const MPSCQueue&) = delete;
    void operator=(const MPSCQueue&) = delete;
};

}} // namespace ppl::common

#endif
