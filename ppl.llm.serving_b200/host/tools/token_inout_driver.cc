// token_inout_driver -- drives the reference's OWN LLMGenerator / LLMEngine / CudaResourceManager (compiled in
// place from the reference tree, linked against libpplnn_b200.so + libb2llm.so) with token-in/out requests,
// i.e. the flow of tools/offline_inference.cc:303-415 without the tokenizer (Request::token_ids set =>
// llm_generator.cc:790-801 bypasses it).  Used by tests/test_host_cpp_gpu.py for end-to-end parity of the C++
// host path against the oracle and by scripts/ for the scheduler steady-state measurement.
//
//   token_inout_driver --model-dir D --model-param-path D/params.json [--quant-method online_i8i8]
//       [--tensor-parallel-size 1] [--max-running-batch 1024] [--max-tokens-per-step 8192]
//       [--max-tokens-scale 0.94] [--top-k 1] [--top-p 0] [--enable-penalty 0] [--enable-profiling 0]
//       (--requests-file F | --requests N --prompt-len P --gen-len G [--seed S]) [--out tokens.txt]
//
// requests file: one request per line "id gen_len tok tok tok ...".  Output file: "id tok tok ..." per request
// (generated tokens in order).  Last stdout line: [RESULT] JSON with the reference's own TPS definition
// (profiler.cc:8-9) and wall-clock tokens/s.
#include "backends/cuda/resource_manager.h"
#include "common/config.h"
#include "common/connection.h"
#include "common/request.h"
#include "common/resource.h"
#include "generator/llm_generator.h"
#include "utils/utils.h"

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>

using namespace ppl::llm;
using namespace ppl::common;

namespace {

class CollectConnection final : public Connection {
public:
    void OnProfiling(const std::shared_ptr<WorkerProfiler>& p) override {
        if (print_profile) PrintProfiler(*p);
        last_profile = p;
    }
    void OnTokenize(uint64_t, const std::vector<int>&) override {}
    void Send(const std::vector<Response>& batch) override {
        std::lock_guard<std::mutex> g(mu);
        for (const auto& r : batch) {
            tokens[r.id].push_back(r.token);
            logprobs[r.id].push_back(r.logprob);
            if (r.finish_flag != FinishFlag::NOT_FINISHED) ++finished;
        }
        if (finished >= wanted) cv.notify_all();
    }
    void NotifyFailure(uint64_t id, RetCode rc, const std::string& msg) override {
        std::lock_guard<std::mutex> g(mu);
        fprintf(stderr, "request %lu failed: %s (%s)\n", (unsigned long)id, msg.c_str(), GetRetCodeStr(rc));
        ++finished;
        ++failed;
        if (finished >= wanted) cv.notify_all();
    }
    void Wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return finished >= wanted; });
    }
    std::mutex mu;
    std::condition_variable cv;
    std::map<uint64_t, std::vector<int>> tokens;
    std::map<uint64_t, std::vector<float>> logprobs;
    std::shared_ptr<WorkerProfiler> last_profile;
    uint64_t finished = 0, wanted = 0, failed = 0;
    bool print_profile = false;
};

std::map<std::string, std::string> ParseArgs(int argc, char** argv) {
    std::map<std::string, std::string> a;
    for (int i = 1; i < argc; ++i) {
        std::string k = argv[i];
        if (k.rfind("--", 0) != 0) {
            fprintf(stderr, "unexpected argument %s\n", argv[i]);
            exit(2);
        }
        const auto eq = k.find('=');
        if (eq != std::string::npos) {
            a[k.substr(2, eq - 2)] = k.substr(eq + 1);
        } else if (i + 1 < argc && strncmp(argv[i + 1], "--", 2) != 0) {
            a[k.substr(2)] = argv[++i];
        } else {
            a[k.substr(2)] = "1";
        }
    }
    return a;
}

} // namespace

int main(int argc, char** argv) {
    auto args = ParseArgs(argc, argv);
    auto gets = [&](const char* k, const char* d) { return args.count(k) ? args[k] : std::string(d); };
    auto geti = [&](const char* k, long d) { return args.count(k) ? atol(args[k].c_str()) : d; };
    auto getf = [&](const char* k, double d) { return args.count(k) ? atof(args[k].c_str()) : d; };

    ResourceConfig resource_config;
    resource_config.model_type = "llama";
    resource_config.model_format = "onnx";
    resource_config.model_dir = gets("model-dir", "");
    resource_config.model_param_path = gets("model-param-path", (resource_config.model_dir + "/params.json").c_str());
    resource_config.tensor_parallel_size = (int32_t)geti("tensor-parallel-size", 1);
    resource_config.max_tokens_scale = (float)getf("max-tokens-scale", 0.94);
    resource_config.max_running_batch = (int32_t)geti("max-running-batch", 1024);
    resource_config.enable_penalty = geti("enable-penalty", 0) != 0;
    resource_config.engine_config.quant_method = gets("quant-method", "online_i8i8");
    resource_config.engine_config.configure_decoding_attn_split_k = (int32_t)geti("configure-decoding-attn-split-k", 1);

    GeneratorConfig generator_config;
    generator_config.top_p = (float)getf("top-p", 0.0);
    generator_config.top_k = (int32_t)geti("top-k", 1);
    generator_config.enable_penalty = resource_config.enable_penalty;
    generator_config.max_running_batch = resource_config.max_running_batch;
    generator_config.max_input_tokens_per_request = (int32_t)geti("max-input-tokens-per-request", 4096);
    generator_config.max_output_tokens_per_request = (int32_t)geti("max-output-tokens-per-request", 4096);
    generator_config.max_total_tokens_per_request = (int32_t)geti("max-total-tokens-per-request", 8192);
    generator_config.max_tokens_per_step = (int32_t)geti("max-tokens-per-step", 8192);
    generator_config.max_cooldown_request = (int)geti("max-cooldown-request", 2);
    generator_config.enable_prefix_cache = geti("enable-prefix-cache", 0) != 0;
    generator_config.max_prefill_batch = (int32_t)geti("max-prefill-batch", generator_config.enable_prefix_cache ? 1 : 64);
    generator_config.enable_profiling = geti("enable-profiling", 0) != 0;

    ModelConfig model_config;
    if (!ParseModelConfig(resource_config.model_param_path, &model_config)) {
        fprintf(stderr, "ParseModelConfig(%s) failed\n", resource_config.model_param_path.c_str());
        return 1;
    }

    // requests
    std::vector<std::shared_ptr<Request>> requests;
    if (args.count("requests-file")) {
        std::ifstream ifs(args["requests-file"]);
        std::string line;
        while (std::getline(ifs, line)) {
            std::istringstream ss(line);
            uint64_t id;
            int gen;
            if (!(ss >> id >> gen)) continue;
            auto r = std::make_shared<Request>();
            r->id = id;
            r->generation_length = gen;
            r->token_ids = std::make_shared<std::vector<int>>();
            int t;
            while (ss >> t) r->token_ids->push_back(t);
            requests.push_back(r);
        }
    } else {
        const long n = geti("requests", 8), plen = geti("prompt-len", 16), glen = geti("gen-len", 8);
        uint64_t s = (uint64_t)geti("seed", 1002) * 0x9E3779B97F4A7C15ull + 1;
        for (long i = 0; i < n; ++i) {
            auto r = std::make_shared<Request>();
            r->id = (uint64_t)i;
            r->generation_length = (int32_t)glen;
            r->token_ids = std::make_shared<std::vector<int>>();
            for (long j = 0; j < plen; ++j) {
                s ^= s << 13; s ^= s >> 7; s ^= s << 17; // xorshift64
                r->token_ids->push_back((int)(s % (uint64_t)model_config.vocab_size));
            }
            requests.push_back(r);
        }
    }
    for (auto& r : requests) {
        r->temperature = (float)getf("temperature", 1.0);
        r->top_k = generator_config.top_k;
        r->top_p = generator_config.top_p;
        r->early_stopping = geti("early-stopping", 0) != 0; // synthetic streams: run to generation_length
        r->repetition_penalty = (float)getf("repetition-penalty", 1.0);
    }

    cuda::CudaResourceManager resource_manager;
    auto rc = resource_manager.Init(model_config, resource_config);
    if (rc != RC_SUCCESS) {
        fprintf(stderr, "CudaResourceManager::Init failed: %s\n", GetRetCodeStr(rc));
        return 1;
    }
    Resource resource;
    resource.tensor_parallel_size = resource_config.tensor_parallel_size;
    resource.kv_cache_max_tokens = resource_manager.kv_cache_max_tokens;
    resource.items = resource_manager.items;
    resource.post_processor = resource_manager.post_processor.get();
    resource.device_worker_pool_ = &resource_manager.device_worker_pool_;
    resource.tokenizer = nullptr; // token-in/out only

    CollectConnection conn;
    conn.wanted = requests.size();
    conn.print_profile = generator_config.enable_profiling;
    auto generator = std::make_unique<LLMGenerator>(resource, generator_config, model_config, &conn);
    rc = generator->Init();
    if (rc != RC_SUCCESS) {
        fprintf(stderr, "LLMGenerator::Init failed: %s\n", GetRetCodeStr(rc));
        return 1;
    }

    uint64_t us = 0;
    {
        utils::TimingGuard t(&us);
        for (auto& r : requests) generator->Process(r);
        conn.Wait();
    }
    uint64_t gen_tokens = 0;
    for (auto& kv : conn.tokens) gen_tokens += kv.second.size();

    if (args.count("out")) {
        std::ofstream ofs(args["out"]);
        for (auto& r : requests) {
            ofs << r->id;
            for (int t : conn.tokens[r->id]) ofs << " " << t;
            ofs << "\n";
        }
    }
    generator.reset();
    printf("[RESULT] {\"requests\": %zu, \"failed\": %lu, \"generated_tokens\": %lu, \"wall_ms\": %.3f, "
           "\"tokens_per_s\": %.1f, \"kv_cache_max_tokens\": %lu}\n",
           requests.size(), (unsigned long)conn.failed, (unsigned long)gen_tokens, us / 1e3,
           us ? gen_tokens * 1e6 / us : 0.0, (unsigned long)resource_manager.kv_cache_max_tokens);
    return conn.failed ? 3 : 0;
}
