// ppl::common::CompactAddrManager (EXTERNAL): first-fit range allocator over an address space that grows on
// demand through a VMAllocator (index_manager.h:27-77: Alloc(nr) -> start or UINTPTR_MAX, Free(start, nr)).
// Free ranges are kept sorted and coalesced, so a freed block merges with its neighbours ("compact").
#ifndef B2LLM_SHIM_PPL_COMMON_COMPACT_ADDR_MANAGER_H_
#define B2LLM_SHIM_PPL_COMMON_COMPACT_ADDR_MANAGER_H_

#include <map>
#include <stdint.h>

namespace ppl { namespace common {

class CompactAddrManager final {
public:
    class VMAllocator {
    public:
        virtual ~VMAllocator() {}
        virtual uintptr_t GetReservedBase() const = 0;
        virtual uint64_t GetAllocatedSize() const = 0;
        /** grow by at least `needed`; returns the amount added, 0 on failure */
        virtual uint64_t Extend(uint64_t needed) = 0;
    };

    explicit CompactAddrManager(VMAllocator* vmr) : vmr_(vmr) {}

    uintptr_t Alloc(uint64_t size) {
        if (size == 0) {
            return UINTPTR_MAX;
        }
        for (auto it = free_.begin(); it != free_.end(); ++it) { // first fit, lowest address
            if (it->second >= size) {
                const uintptr_t start = it->first;
                const uint64_t rest = it->second - size;
                free_.erase(it);
                if (rest) {
                    free_[start + size] = rest;
                }
                return start;
            }
        }
        // grow: a trailing free block is extended rather than wasted
        const uintptr_t end = vmr_->GetReservedBase() + vmr_->GetAllocatedSize();
        uint64_t have = 0;
        uintptr_t start = end;
        if (!free_.empty()) {
            auto last = std::prev(free_.end());
            if (last->first + last->second == end) {
                have = last->second;
                start = last->first;
            }
        }
        const uint64_t got = vmr_->Extend(size - have);
        if (got < size - have) {
            if (got) {
                Free(end, got);
            }
            return UINTPTR_MAX;
        }
        if (have) {
            free_.erase(start);
        }
        if (have + got > size) {
            free_[start + size] = have + got - size;
        }
        return start;
    }

    void Free(uintptr_t start, uint64_t size) {
        if (size == 0) {
            return;
        }
        auto next = free_.lower_bound(start);
        if (next != free_.begin()) {
            auto prev = std::prev(next);
            if (prev->first + prev->second == start) {
                start = prev->first;
                size += prev->second;
                free_.erase(prev);
            }
        }
        if (next != free_.end() && start + size == next->first) {
            size += next->second;
            free_.erase(next);
        }
        free_[start] = size;
    }

private:
    VMAllocator* vmr_;
    std::map<uintptr_t, uint64_t> free_; // start -> length
};

}} // namespace ppl::common

#endif
