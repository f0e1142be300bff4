// ppl::common::TypedMPSCQueue<T> (EXTERNAL): value queue over MPSCQueue
// (llm_generator.h:107-119, llm_generator.cc:392,729,750: Push(T), bool Pop(T*), Size()).
#ifndef B2LLM_SHIM_PPL_COMMON_TYPED_MPSC_QUEUE_H_
#define B2LLM_SHIM_PPL_COMMON_TYPED_MPSC_QUEUE_H_

#include "mpsc_queue.h"

#include <stdint.h>
#include <utility>

namespace ppl { namespace common {

template <typename T>
class TypedMPSCQueue final {
public:
    ~TypedMPSCQueue() {
        T tmp;
        while (Pop(&tmp)) {
        }
    }
    void Push(const T& v) {
        auto* n = new Item(v);
        queue_.Push(n);
        size_.fetch_add(1, std::memory_order_acq_rel);
    }
    void Push(T&& v) {
        auto* n = new Item(std::move(v));
        queue_.Push(n);
        size_.fetch_add(1, std::memory_order_acq_rel);
    }
    bool Pop(T* out) {
        bool is_empty = true;
        MPSCQueue::Node* node;
        do {
            node = queue_.Pop(&is_empty);
        } while (!node && !is_empty);
        if (!node) {
            return false;
        }
        auto* item = static_cast<Item*>(node);
        *out = std::move(item->value);
        delete item;
        size_.fetch_sub(1, std::memory_order_acq_rel);
        return true;
    }
    uint32_t Size() const {
        return size_.load(std::memory_order_acquire);
    }

private:
    struct Item final : public MPSCQueue::Node {
        explicit Item(const T& v) : value(v) {}
        explicit Item(T&& v) : value(std::move(v)) {}
        T value;
    };
    MPSCQueue queue_;
    std::atomic<uint32_t> size_{0};
};

}} // namespace ppl::common

#endif
