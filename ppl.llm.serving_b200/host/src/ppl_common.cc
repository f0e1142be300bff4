// Out-of-line parts of the ppl::common stand-in (host/include/ppl/common/*): logging, RetCode names, the
// static thread pool, CUDA / NCCL environment set-up.  ppl.common itself is external to the reference tree
// (hpcc @ master, cmake/deps.cmake:32-58); the reference's call sites define what is needed.
#include "ppl/common/cuda/cuda_env.h"
#include "ppl/common/cuda/nccl_utils.h"
#include "ppl/common/log.h"
#include "ppl/common/retcode.h"
#include "ppl/common/threadpool.h"
#include "ppl/common/types.h"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>

namespace ppl { namespace common {

// ---------------------------------------------------------------------------------------------- RetCode
const char* GetRetCodeStr(RetCode rc) {
    static const char* names[] = {"success",       "other error",      "invalid value",       "out of memory",
                                  "unsupported",   "device runtime error", "device memory error", "not found",
                                  "exists",        "out of range",     "permission denied",   "sign in"};
    return rc < sizeof(names) / sizeof(names[0]) ? names[rc] : "unknown retcode";
}

const char* GetDataTypeStr(datatype_t dt) {
    static const char* names[] = {"UNKNOWN", "UINT8",    "UINT16", "UINT32", "UINT64", "FLOAT16", "FLOAT32", "FLOAT64",
                                  "BFLOAT16", "INT4B",   "INT8",   "INT16",  "INT32",  "INT64",   "BOOL"};
    return dt < DATATYPE_MAX ? names[dt] : "UNKNOWN";
}

// ---------------------------------------------------------------------------------------------- logging
namespace {
int InitialLogLevel() {
    const char* s = getenv("PPL_LOG_LEVEL");
    if (!s) return LOG_LEVEL_INFO;
    if (!strcasecmp(s, "DEBUG")) return LOG_LEVEL_DEBUG;
    if (!strcasecmp(s, "WARNING")) return LOG_LEVEL_WARNING;
    if (!strcasecmp(s, "ERROR")) return LOG_LEVEL_ERROR;
    return LOG_LEVEL_INFO;
}
int g_log_level = InitialLogLevel();
std::mutex g_log_mutex;
} // namespace

int GetCurrentLogLevel() {
    return g_log_level;
}
void SetCurrentLogLevel(int l) {
    g_log_level = l;
}

LogMessage::LogMessage(int level, const char* file, int line) : enabled_(level >= g_log_level), level_(level) {
    if (!enabled_) return;
    static const char* tags[] = {"DEBUG", "INFO", "WARNING", "ERROR", "FATAL"};
    const char* base = strrchr(file, '/');
    const auto now = std::chrono::system_clock::now();
    const time_t t = std::chrono::system_clock::to_time_t(now);
    const int ms = (int)(std::chrono::duration_cast<std::chrono::milliseconds>(now.time_since_epoch()).count() % 1000);
    struct tm tmv;
    localtime_r(&t, &tmv);
    char head[96];
    snprintf(head, sizeof(head), "[%s][%04d-%02d-%02d %02d:%02d:%02d.%03d][%s:%d] ", tags[level], tmv.tm_year + 1900,
             tmv.tm_mon + 1, tmv.tm_mday, tmv.tm_hour, tmv.tm_min, tmv.tm_sec, ms, base ? base + 1 : file, line);
    os_ << head;
}

LogMessage::~LogMessage() {
    if (!enabled_) return;
    os_ << '\n';
    const std::string s = os_.str();
    std::lock_guard<std::mutex> g(g_log_mutex);
    fwrite(s.data(), 1, s.size(), stderr);
    if (level_ >= LOG_LEVEL_ERROR) fflush(stderr);
}

// ---------------------------------------------------------------------------------------------- thread pool
StaticThreadPool::~StaticThreadPool() {
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        stop_ = true;
    }
    cv_start_.notify_all();
    for (auto& t : workers_) {
        if (t.joinable()) t.join();
    }
}

RetCode StaticThreadPool::Init(uint32_t thread_num) {
    if (!workers_.empty() || thread_num == 0) {
        return RC_INVALID_VALUE;
    }
    workers_.reserve(thread_num);
    for (uint32_t i = 0; i < thread_num; ++i) {
        workers_.emplace_back([this, i] { WorkerLoop(i); });
    }
    return RC_SUCCESS;
}

void StaticThreadPool::WorkerLoop(uint32_t ithr) {
    uint64_t seen = 0;
    const uint32_t nthr = (uint32_t)workers_.capacity();
    while (true) {
        std::function<void(uint32_t, uint32_t)> task;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_start_.wait(lk, [&] { return stop_ || round_ != seen; });
            if (stop_) return;
            seen = round_;
            task = task_;
        }
        task(nthr, ithr);
        {
            std::lock_guard<std::mutex> g(mu_);
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }
}

void StaticThreadPool::RunAsync(const std::function<void(uint32_t, uint32_t)>& f) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
    task_ = f;
    pending_ = (uint32_t)workers_.size();
    ++round_;
    lk.unlock();
    cv_start_.notify_all();
}

void StaticThreadPool::Wait() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
}

void StaticThreadPool::Run(const std::function<void(uint32_t, uint32_t)>& f) {
    RunAsync(f);
    Wait();
}

// ---------------------------------------------------------------------------------------------- CUDA / NCCL
RetCode InitCudaEnv(int device_id) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        LOG(ERROR) << "InitCudaEnv: no CUDA device (b2llm has no CPU path)";
        return RC_DEVICE_RUNTIME_ERROR;
    }
    if (device_id < 0 || device_id >= n) {
        LOG(ERROR) << "InitCudaEnv: device [" << device_id << "] out of range, " << n << " device(s) visible";
        return RC_INVALID_VALUE;
    }
    const cudaError_t err = cudaSetDevice(device_id);
    if (err != cudaSuccess) {
        LOG(ERROR) << "cudaSetDevice(" << device_id << ") failed: " << cudaGetErrorString(err);
        return RC_DEVICE_RUNTIME_ERROR;
    }
    cudaFree(nullptr); // create the primary context on this thread now
    return RC_SUCCESS;
}

#ifdef PPLNN_CUDA_ENABLE_NCCL
RetCode InitNccl(uint32_t tensor_parallel_size, std::vector<ncclComm_t>* nccl_comm_list) {
    nccl_comm_list->assign(tensor_parallel_size, nullptr);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        if (tensor_parallel_size <= 1) {
            return RC_SUCCESS; // no communicator needed; the missing device is reported by InitCudaEnv
        }
        LOG(ERROR) << "InitNccl: no CUDA device";
        return RC_DEVICE_RUNTIME_ERROR;
    }
    if ((int)tensor_parallel_size > ndev) {
        LOG(ERROR) << "InitNccl: tensor_parallel_size " << tensor_parallel_size << " > " << ndev << " visible device(s)";
        return RC_INVALID_VALUE;
    }
    std::vector<int> devs(tensor_parallel_size);
    for (uint32_t i = 0; i < tensor_parallel_size; ++i) devs[i] = (int)i;
    const ncclResult_t r = ncclCommInitAll(nccl_comm_list->data(), (int)tensor_parallel_size, devs.data());
    if (r != ncclSuccess) {
        LOG(ERROR) << "ncclCommInitAll failed: " << ncclGetErrorString(r);
        nccl_comm_list->clear();
        return RC_OTHER_ERROR;
    }
    return RC_SUCCESS;
}
#endif

}} // namespace ppl::common
