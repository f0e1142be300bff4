// CPU unit tests of the ppl::common stand-in (host/include/ppl/common/*): the containers and primitives the
// reference's generator relies on.  No GPU needed.  Exit code 0 = all passed; prints one line per case.
#include "ppl/common/barrier.h"
#include "ppl/common/compact_addr_manager.h"
#include "ppl/common/event_count.h"
#include "ppl/common/log.h"
#include "ppl/common/mpsc_queue.h"
#include "ppl/common/page_manager.h"
#include "ppl/common/retcode.h"
#include "ppl/common/threadpool.h"
#include "ppl/common/typed_mpsc_queue.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <set>
#include <thread>
#include <vector>

using namespace ppl::common;

static int g_failed = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++g_failed;                                                      \
        }                                                                    \
    } while (0)

// the allocator the reference builds on CompactAddrManager (src/utils/index_manager.h:27-50), restated
class TestVM final : public CompactAddrManager::VMAllocator {
public:
    explicit TestVM(uint64_t max) : max_(max) {}
    uintptr_t GetReservedBase() const override { return 0; }
    uint64_t GetAllocatedSize() const override { return used_; }
    uint64_t Extend(uint64_t needed) override {
        if (needed + used_ > max_) return 0;
        used_ += needed;
        return needed;
    }

private:
    uint64_t max_, used_ = 0;
};

static void test_page_manager() {
    PageManager pm;
    pm.Init(100, 16);  // 6 whole pages
    CHECK(pm.GetAvail() == 6);
    std::vector<int64_t> a, b;
    CHECK(pm.Alloc(4, &a) == RC_SUCCESS && a.size() == 4);
    for (auto p : a) CHECK(p % 16 == 0 && p >= 0 && p < 96);  // entries are first token slots of pages
    CHECK(std::set<int64_t>(a.begin(), a.end()).size() == 4);
    CHECK(pm.Alloc(3, &b) != RC_SUCCESS && b.empty() && pm.GetAvail() == 2);  // failure has no side effects
    CHECK(pm.Alloc(2, &b) == RC_SUCCESS && pm.GetAvail() == 0);
    pm.Free(a.data(), a.size());
    CHECK(pm.GetAvail() == 4);
    std::vector<int64_t> c = {7};
    CHECK(pm.Alloc(1, &c) == RC_SUCCESS && c.size() == 2 && c[0] == 7);  // appends
    printf("page_manager ok\n");
}

static void test_compact_addr_manager() {
    TestVM vm(100);
    CompactAddrManager m(&vm);
    const uintptr_t a = m.Alloc(30), b = m.Alloc(30), c = m.Alloc(30);
    CHECK(a == 0 && b == 30 && c == 60);
    CHECK(m.Alloc(20) == UINTPTR_MAX);  // 10 left
    m.Free(b, 30);
    CHECK(m.Alloc(40) == UINTPTR_MAX);  // hole of 30 + tail of 10 are not adjacent
    const uintptr_t d = m.Alloc(20);
    CHECK(d == 30);                     // first fit, lowest address
    m.Free(c, 30);                      // [50,60) free + [60,90) free + tail -> coalesced with the hole's rest
    const uintptr_t e = m.Alloc(50);    // [50, 100): needs the rest of the hole, c's block and the unextended tail
    CHECK(e == 50);
    m.Free(a, 30);
    m.Free(d, 20);
    m.Free(e, 50);
    CHECK(m.Alloc(100) == 0);           // everything merged back into one range
    printf("compact_addr_manager ok\n");
}

struct Node final : public MPSCQueue::Node {
    explicit Node(int v) : value(v) {}
    int value;
};

static void test_mpsc_queues() {
    MPSCQueue q;
    bool empty = false;
    CHECK(q.Pop(&empty) == nullptr && empty);
    constexpr int kProducers = 4, kPer = 5000;
    std::vector<std::thread> ths;
    for (int p = 0; p < kProducers; ++p)
        ths.emplace_back([&q, p] {
            for (int i = 0; i < kPer; ++i) q.Push(new Node(p * kPer + i));
        });
    std::vector<int> last(kProducers, -1);
    int got = 0;
    while (got < kProducers * kPer) {
        auto* n = q.Pop(&empty);
        if (!n) continue;
        auto* node = static_cast<Node*>(n);
        const int p = node->value / kPer, i = node->value % kPer;
        CHECK(i == last[p] + 1);  // per-producer FIFO
        last[p] = i;
        delete node;
        ++got;
    }
    for (auto& t : ths) t.join();
    CHECK(q.Pop(&empty) == nullptr && empty);

    TypedMPSCQueue<int> tq;
    int v = 0;
    CHECK(!tq.Pop(&v) && tq.Size() == 0);
    tq.Push(3);
    tq.Push(4);
    CHECK(tq.Size() == 2 && tq.Pop(&v) && v == 3 && tq.Pop(&v) && v == 4 && !tq.Pop(&v));
    printf("mpsc queues ok\n");
}

static void test_threadpool_and_barrier() {
    StaticThreadPool pool;
    CHECK(pool.Init(4) == RC_SUCCESS && pool.GetNumThreads() == 4);
    // thread i always handles index i (the reference binds CUDA devices this way)
    std::vector<std::thread::id> ids(4);
    pool.Run([&](uint32_t nthr, uint32_t ithr) {
        CHECK(nthr == 4);
        ids[ithr] = std::this_thread::get_id();
    });
    for (int round = 0; round < 20; ++round)
        pool.Run([&](uint32_t, uint32_t ithr) { CHECK(ids[ithr] == std::this_thread::get_id()); });
    std::atomic<int> n{0};
    pool.RunAsync([&](uint32_t, uint32_t) {
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
        ++n;
    });
    pool.Wait();
    CHECK(n == 4);
    Barrier bar;
    bar.Reset(4);
    std::atomic<int> before{0}, after_min{100};
    pool.Run([&](uint32_t, uint32_t) {
        ++before;
        bar.Wait();
        int b = before.load();
        int cur = after_min.load();
        while (b < cur && !after_min.compare_exchange_weak(cur, b)) {
        }
    });
    CHECK(after_min == 4);  // nobody passed the barrier before all four arrived
    printf("threadpool + barrier ok\n");
}

static void test_event_count() {
    // the generator thread's sleep protocol (llm_generator.cc:342-366): a notification between PrepareWait and
    // CommitWait must not be lost
    EventCount ec;
    std::atomic<int> pending{0}, consumed{0};
    std::atomic<bool> stop{false};
    std::thread consumer([&] {
        while (true) {
            if (pending.load() > 0) {
                --pending;
                ++consumed;
                continue;
            }
            if (stop.load()) break;
            auto key = ec.PrepareWait();
            if (pending.load() > 0 || stop.load()) {
                ec.CancelWait();
                continue;
            }
            ec.CommitWait(key);
        }
    });
    for (int i = 0; i < 2000; ++i) {
        ++pending;
        ec.NotifyOne();
        if (i % 64 == 0) std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    while (consumed.load() < 2000) std::this_thread::yield();
    stop = true;
    ec.NotifyOne();
    consumer.join();
    CHECK(consumed == 2000);
    printf("event_count ok\n");
}

int main() {
    SetCurrentLogLevel(LOG_LEVEL_ERROR);
    CHECK(std::string(GetRetCodeStr(RC_SUCCESS)) == "success" && std::string(GetRetCodeStr(RC_INVALID_VALUE)) == "invalid value");
    test_page_manager();
    test_compact_addr_manager();
    test_mpsc_queues();
    test_threadpool_and_barrier();
    test_event_count();
    printf(g_failed ? "FAILED: %d checks\n" : "all ok (%d failures)\n", g_failed);
    return g_failed ? 1 : 0;
}
