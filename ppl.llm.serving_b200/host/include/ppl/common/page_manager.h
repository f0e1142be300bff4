// ppl::common::PageManager (EXTERNAL): KV page allocator of cache_mode 1
// (llm_generator.cc:157 Init(max_tokens, page_size); :534,554 Alloc(n, &page_list) -> RetCode;
//  :376,423-425,522 Free(ptr, n); :516,775 GetAvail() in pages).
// Entries handed out are the FIRST TOKEN SLOT of a page (page_id * page_size): the generator copies them
// verbatim into ModelInput::page_list (llm_generator.cc:293-296) and the attention / KV-append kernels
// address token p of a sequence at page_list[b][p / page_size] + p % page_size.
#ifndef B2LLM_SHIM_PPL_COMMON_PAGE_MANAGER_H_
#define B2LLM_SHIM_PPL_COMMON_PAGE_MANAGER_H_

#include "retcode.h"

#include <stdint.h>
#include <vector>

namespace ppl { namespace common {

class PageManager final {
public:
    void Init(uint64_t max_tokens, uint64_t page_size) {
        page_size_ = page_size ? page_size : 1;
        const uint64_t n = max_tokens / page_size_;
        free_.clear();
        free_.reserve(n);
        for (uint64_t i = n; i > 0; --i) { // pop_back hands out low pages first
            free_.push_back((int64_t)((i - 1) * page_size_));
        }
    }
    /** appends n pages to *page_list, or fails without side effects */
    RetCode Alloc(uint64_t n, std::vector<int64_t>* page_list) {
        if (n > free_.size()) {
            return RC_OUT_OF_MEMORY;
        }
        for (uint64_t i = 0; i < n; ++i) {
            page_list->push_back(free_.back());
            free_.pop_back();
        }
        return RC_SUCCESS;
    }
    void Free(const int64_t* pages, uint64_t n) {
        for (uint64_t i = 0; i < n; ++i) {
            free_.push_back(pages[i]);
        }
    }
    int64_t GetAvail() const {
        return (int64_t)free_.size();
    }

private:
    uint64_t page_size_ = 1;
    std::vector<int64_t> free_;
};

}} // namespace ppl::common

#endif
