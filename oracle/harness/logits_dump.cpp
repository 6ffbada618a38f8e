// logits_dump.cpp -- TEST INFRASTRUCTURE.  Drives the reference's public API (llama.h: llama_decode) exactly like
// LlamaServerContext does (C/src/llama_server_context.cc:1635) on a fixed, seeded token sequence and dumps every
// logit row, so the CPU backend and the B200 backend can be compared on identical inputs (teacher forcing).
//   logits_dump MODEL OUT.bin NGL N_PROMPT N_GEN [kv_type f16|q8_0|q4_0] [n_parallel] [threads]
// Output: int32 n_rows, int32 n_vocab, then n_rows*n_vocab f32.  Also prints per-phase timings.
// OUT.bin == "-" is bench mode (bench.py e2e / --impl reference legs): nothing is stored, the prompt only asks for the
// logits of its last token, and env LOGITS_DUMP_WARMUP=W runs W untimed decode steps before the n_gen timed ones.
// env LOGITS_DUMP_GREEDY=1: instead of teacher forcing, every slot feeds back the argmax of its previous logits row
// (llama-cli --temp 0 --top-k 1); the chosen token ids are written to OUT.bin.tok (int32 [n_gen][n_parallel]) so two
// backends' greedy streams can be compared token for token.  env LOGITS_DUMP_SEED=S reseeds the synthetic prompt.
#include "llama.h"
#include "ggml.h"
#include "ggml-backend.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// optional per-node dump (env LOGITS_DUMP_NODES=file): every f32 node output of every graph, in execution order,
// through llama's cb_eval hook (llama-context.cpp:1350) -- lets tools/compare_nodes.py find the first op that deviates
static FILE *g_nodes = nullptr;
static bool node_cb(struct ggml_tensor *t, bool ask, void *) {
    if (ask) return t->type == GGML_TYPE_F32;
    const int64_t n = ggml_nelements(t);
    if (!ggml_is_contiguous(t) || n > (1 << 22)) return true;
    std::vector<float> buf((size_t)n);
    ggml_backend_tensor_get(t, buf.data(), 0, (size_t)n * 4);
    char name[96] = {0};
    snprintf(name, sizeof(name), "%s|%s", t->name, ggml_op_desc(t));
    fwrite(name, 1, sizeof(name), g_nodes);
    fwrite(&n, 8, 1, g_nodes);
    fwrite(buf.data(), 4, (size_t)n, g_nodes);
    return true;
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s model out ngl n_prompt n_gen [kv] [n_parallel] [threads]\n", argv[0]); return 2; }
    const char *model_path = argv[1], *out_path = argv[2];
    const int ngl = atoi(argv[3]), n_prompt = atoi(argv[4]), n_gen = atoi(argv[5]);
    const char *kv = argc > 6 ? argv[6] : "f16";
    const int n_par = argc > 7 ? atoi(argv[7]) : 1;
    const int threads = argc > 8 ? atoi(argv[8]) : 8;
    const bool bench = !strcmp(out_path, "-");
    const int n_warm = bench && getenv("LOGITS_DUMP_WARMUP") ? atoi(getenv("LOGITS_DUMP_WARMUP")) : 0;
    ggml_backend_load_all();
    llama_backend_init();
    llama_model_params mp = llama_model_default_params();
    mp.n_gpu_layers = ngl;
    if (const char *sm = getenv("LOGITS_DUMP_SPLIT_MODE")) {      // "row": weights row-split over all GPUs (ggml_backend_split_buffer_type), "none": main GPU only
        if (!strcmp(sm, "row")) mp.split_mode = LLAMA_SPLIT_MODE_ROW;
        else if (!strcmp(sm, "none")) mp.split_mode = LLAMA_SPLIT_MODE_NONE;
    }
    llama_model *model = llama_model_load_from_file(model_path, mp);
    if (!model) return 1;
    llama_context_params cp = llama_context_default_params();
    cp.n_ctx = (n_prompt + n_warm + n_gen + 8) * n_par;
    cp.n_batch = 2048; cp.n_ubatch = 2048;                 // cortex defaults (C/src/llama_engine.cc:617-618)
    cp.n_seq_max = n_par;
    cp.flash_attn = true;
    cp.n_threads = threads; cp.n_threads_batch = threads;
    cp.type_k = cp.type_v = !strcmp(kv, "q8_0") ? GGML_TYPE_Q8_0 : !strcmp(kv, "q4_0") ? GGML_TYPE_Q4_0 : GGML_TYPE_F16;
    if (const char *np = getenv("LOGITS_DUMP_NODES")) { g_nodes = fopen(np, "wb"); cp.cb_eval = node_cb; cp.cb_eval_user_data = nullptr; }
    llama_context *ctx = llama_init_from_model(model, cp);
    if (!ctx) return 1;
    const int V = llama_vocab_n_tokens(llama_model_get_vocab(model));
    // seeded token ids in [3, V)
    uint64_t s = 0x9E3779B97F4A7C15ull;
    if (const char *sd = getenv("LOGITS_DUMP_SEED")) s ^= (uint64_t)atoll(sd) * 0xD1B54A32D192ED03ull;
    const bool greedy = getenv("LOGITS_DUMP_GREEDY") && atoi(getenv("LOGITS_DUMP_GREEDY")) != 0;
    std::vector<llama_token> last_tok((size_t)n_par, 0);
    std::vector<int32_t> stream;
    auto argmax = [&](const float *l) { int b = 0; for (int i = 1; i < V; i++) if (l[i] > l[b]) b = i; return (llama_token)b; };
    auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (llama_token)(3 + (s % (uint64_t)(V - 3))); };
    std::vector<float> all;
    int rows = 0;
    llama_batch batch = llama_batch_init(n_prompt * n_par > n_par ? n_prompt * n_par : n_par, 0, 1);
    // ---- prefill: every slot gets its own prompt, all in one batch (continuous batching shape) ----
    batch.n_tokens = 0;
    for (int p = 0; p < n_par; p++)
        for (int i = 0; i < n_prompt; i++) {
            const int j = batch.n_tokens++;
            batch.token[j] = next(); batch.pos[j] = i; batch.n_seq_id[j] = 1; batch.seq_id[j][0] = p; batch.logits[j] = !bench || i == n_prompt - 1;
        }
    if (bench && n_warm > 0) {            // like llama-bench: one untimed pass first (module load, first-use allocations), then an empty cache again
        if (llama_decode(ctx, batch) != 0) { fprintf(stderr, "prefill warm-up failed\n"); return 1; }
        llama_synchronize(ctx);
        llama_kv_self_clear(ctx);
    }
    double t0 = now_ms();
    if (llama_decode(ctx, batch) != 0) { fprintf(stderr, "prefill decode failed\n"); return 1; }
    llama_synchronize(ctx);
    const double t_prefill = now_ms() - t0;
    if (!bench) for (int j = 0; j < batch.n_tokens; j++) { const float *l = llama_get_logits_ith(ctx, j); all.insert(all.end(), l, l + V); rows++; }
    if (greedy) for (int p = 0; p < n_par; p++) last_tok[p] = argmax(llama_get_logits_ith(ctx, p * n_prompt + n_prompt - 1));
    // ---- decode: one token per slot per step, teacher forced ----
    t0 = now_ms();
    float sink = 0.0f;
    for (int g = -n_warm; g < n_gen; g++) {
        if (g == 0) { llama_synchronize(ctx); t0 = now_ms(); }
        batch.n_tokens = 0;
        for (int p = 0; p < n_par; p++) {
            const int j = batch.n_tokens++;
            batch.token[j] = greedy ? last_tok[p] : next(); batch.pos[j] = n_prompt + n_warm + g;
            if (greedy) stream.push_back(batch.token[j]); batch.n_seq_id[j] = 1; batch.seq_id[j][0] = p; batch.logits[j] = 1;
        }
        if (llama_decode(ctx, batch) != 0) { fprintf(stderr, "decode failed at %d\n", g); return 1; }
        // reading the logits is what the sampler does every step: it forces the device->host copy + synchronize
        for (int j = 0; j < batch.n_tokens; j++) {
            const float *l = llama_get_logits_ith(ctx, j);
            if (bench) sink += l[0] + l[V - 1]; else { all.insert(all.end(), l, l + V); rows++; }
            if (greedy) last_tok[j] = argmax(l);
        }
    }
    llama_synchronize(ctx);
    const double t_decode = now_ms() - t0;
    if (!bench) {
        FILE *f = fopen(out_path, "wb");
        fwrite(&rows, 4, 1, f); fwrite(&V, 4, 1, f); fwrite(all.data(), 4, all.size(), f); fclose(f);
        if (greedy) {
            std::string tp = std::string(out_path) + ".tok";
            FILE *ft = fopen(tp.c_str(), "wb");
            fwrite(stream.data(), 4, stream.size(), ft); fclose(ft);
        }
    } else if (sink != sink) fprintf(stderr, "nan in logits\n");
    printf("{\"ngl\": %d, \"n_parallel\": %d, \"n_prompt\": %d, \"n_gen\": %d, \"kv\": \"%s\", \"prefill_ms\": %.3f, \"prefill_tok_s\": %.1f, \"decode_ms\": %.3f, \"decode_tok_s\": %.1f}\n",
           ngl, n_par, n_prompt, n_gen, kv, t_prefill, 1000.0 * n_prompt * n_par / t_prefill, t_decode, n_gen > 0 ? 1000.0 * n_gen * n_par / t_decode : 0.0);
    llama_batch_free(batch); llama_free(ctx); llama_model_free(model);
    return 0;
}
