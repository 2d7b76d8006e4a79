// wva_kernels.cuh — sm_100a kernels of the Analyze -> Optimize path.
//
//   k_pairs          one thread per (server, accelerator): core.CreateAllocation + penalty
//   k_grid           candidate sweep: one thread per (server, accel, replicas, batch) candidate,
//                    service-rate tables staged in shared memory, warp-shuffle + block argmin
//   k_grid_list      deferred long chains / materialised-p[] chains, one thread per candidate
//   k_grid_claim     per-server winners from the published (key, metrics) slots
//   k_solve_unlimited / k_greedy_*   the assignment step
//   k_totals         System.AllocateByType partial sums (input of the one NCCL allreduce)
//   k_queue_analyze / k_queue_size   the pkg/analyzer public API, batched
#pragma once

#include "wva_device.cuh"

namespace wva {

// Device copy of wva_system_soa (device pointers).
struct DevSystem {
    int S, A, M, T;
    int cert;      // tuning: certified closed-form tails on (1) / off (0); results identical either way
    const float* acc_cost; const int* acc_multiplicity; const int* acc_type;
    const long long* type_capacity;
    const float *perf_alpha, *perf_beta, *perf_gamma, *perf_delta;
    const int *perf_max_batch, *perf_at_tokens, *perf_acc_count; const unsigned char* perf_valid;
    const int* srv_model; const float* srv_arrival_rpm; const int *srv_in_tokens, *srv_out_tokens;
    const float *srv_slo_ttft, *srv_slo_itl, *srv_slo_tps; const unsigned char* srv_target_valid;
    const int *srv_priority, *srv_min_replicas, *srv_max_batch; const unsigned char* srv_keep_acc;
    const int *srv_cur_acc, *srv_cur_replicas; const float* srv_cur_cost;
};

struct DevAllocs {   // wva_alloc_soa with device pointers
    int* acc; long long* num_replicas; long long* batch_size;
    float *cost, *value, *itl, *ttft, *rho, *max_arrv;
};

__device__ __forceinline__ void store_alloc(const DevAllocs& o, size_t i, const AllocRec& a) {
    o.acc[i] = a.acc; o.num_replicas[i] = a.numReplicas; o.batch_size[i] = a.batchSize;
    o.cost[i] = a.cost; o.value[i] = a.value; o.itl[i] = a.itl; o.ttft[i] = a.ttft; o.rho[i] = a.rho;
    o.max_arrv[i] = a.maxArrv;
}
__device__ __forceinline__ AllocRec load_alloc(const DevAllocs& o, size_t i) {
    AllocRec a;
    a.acc = o.acc[i]; a.numReplicas = o.num_replicas[i]; a.batchSize = o.batch_size[i];
    a.cost = o.cost[i]; a.value = o.value[i]; a.itl = o.itl[i]; a.ttft = o.ttft[i]; a.rho = o.rho[i];
    a.maxArrv = o.max_arrv[i];
    return a;
}

__device__ __forceinline__ long long num_instances(const DevSystem& sys, int m, int a) {   // model.go:45-54
    int c = sys.perf_acc_count[(size_t)m * sys.A + a];
    return c <= 0 ? 1 : c;
}
// the nil-returning lookups of CreateAllocation, allocation.go:41-70
__device__ __forceinline__ bool pair_lookups_ok(const DevSystem& sys, int s, int a) {
    if (sys.srv_arrival_rpm[s] < 0.0f || sys.srv_in_tokens[s] < 0 || sys.srv_out_tokens[s] < 0) return false;
    int m = sys.srv_model[s];
    if (m < 0 || m >= sys.M) return false;
    if (!sys.perf_valid[(size_t)m * sys.A + a]) return false;
    if (!sys.srv_target_valid[s]) return false;
    return true;
}
// Server.GetCandidateAccelerators, server.go:70-82
__device__ __forceinline__ bool is_candidate_accel(const DevSystem& sys, int s, int a) {
    if (sys.srv_keep_acc[s]) {
        int cur = sys.srv_cur_acc[s];
        if (cur != WVA_ACC_NONE) return cur == a;
    }
    return true;
}

// zeroLoadAllocation, allocation.go:259-288
__device__ AllocRec zero_load_allocation(const DevSystem& sys, int s, int a) {
    AllocRec out = empty_alloc();
    long long numReplicas = sys.srv_min_replicas[s];
    if (numReplicas == 0) return out;
    int m = sys.srv_model[s];
    size_t pi = (size_t)m * sys.A + a;
    long long maxBatch = sys.perf_max_batch[pi];
    if (sys.srv_max_batch[s] > 0) maxBatch = sys.srv_max_batch[s];
    long long total = go_muli(num_instances(sys, m, a), numReplicas);
    float cost = sys.acc_cost[a] * (float)total;
    float alpha = sys.perf_alpha[pi], beta = sys.perf_beta[pi], gamma = sys.perf_gamma[pi], delta = sys.perf_delta[pi];
    float decode = alpha + beta;
    float bb = beta * (float)maxBatch;
    float maxDecode = alpha + bb;
    float prefill = gamma + delta;
    float maxServ = prefill + maxDecode;
    out.acc = a; out.numReplicas = numReplicas; out.batchSize = maxBatch;
    out.cost = cost; out.itl = decode; out.ttft = prefill; out.rho = 0.0f;
    out.maxArrv = (float)maxBatch / maxServ;
    out.value = cost;
    return out;
}

// N of CreateAllocation (allocation.go:77-87); 0 when the pair never reaches the queue analyzer.
__device__ __forceinline__ long long pair_batch_size(const DevSystem& sys, int s, int a) {
    if (!pair_lookups_ok(sys, s, a) || !is_candidate_accel(sys, s, a)) return 0;
    if (sys.srv_arrival_rpm[s] == 0.0f || sys.srv_out_tokens[s] == 0) return 0;
    if (sys.srv_max_batch[s] > 0) return sys.srv_max_batch[s];
    size_t pi = (size_t)sys.srv_model[s] * sys.A + a;
    long long n = go_divi(go_muli(sys.perf_max_batch[pi], sys.perf_at_tokens[pi]), sys.srv_out_tokens[s]);
    return n > 1 ? n : 1;
}

// core.CreateAllocation, allocation.go:27-163.  Returns false for nil.  `fault` is set when the
// streaming solver met an overflow-rescale case and no scratch was supplied (caller re-runs the
// pair in the literal kernel).
__device__ bool create_allocation(const DevSystem& sys, int s, int a, double* scratch, AllocRec& out, int& fault,
                                  unsigned long long& steps) {
    fault = 0;
    if (!pair_lookups_ok(sys, s, a)) return false;
    const float arrival = sys.srv_arrival_rpm[s];
    const long long inTok = sys.srv_in_tokens[s], outTok = sys.srv_out_tokens[s];
    if (arrival == 0.0f || outTok == 0) { out = zero_load_allocation(sys, s, a); return true; }
    const int m = sys.srv_model[s];
    const size_t pi = (size_t)m * sys.A + a;
    const long long K = outTok;
    long long N;
    if (sys.srv_max_batch[s] > 0) N = sys.srv_max_batch[s];
    else { N = go_divi(go_muli(sys.perf_max_batch[pi], sys.perf_at_tokens[pi]), K); if (N < 1) N = 1; }
    const long long maxQueue = go_muli(N, WVA_MAX_QUEUE_TO_BATCH_RATIO);
    if (!config_ok(N, maxQueue, inTok, K)) return false;
    ServiceParms sp; sp.alpha = sys.perf_alpha[pi]; sp.beta = sys.perf_beta[pi];
    sp.gamma = sys.perf_gamma[pi]; sp.delta = sys.perf_delta[pi];
    Analyzer qa;
    qa.build(sp, N, maxQueue, inTok, K, scratch);
    qa.cert = sys.cert != 0;
    const float tTTFT = sys.srv_slo_ttft[s], tITL = sys.srv_slo_itl[s], tTPS = sys.srv_slo_tps[s];
    float rates[3], achieved[3];
    wva_metrics metrics;
    bool ok = size_queue(qa, tTTFT, tITL, tTPS, rates, metrics, achieved);
    if (qa.fault) { fault = qa.fault; steps += qa.steps; return false; }
    if (!ok) { steps += qa.steps; return false; }
    const float rateStar = metrics.throughput;
    float totalRate;
    if (tTPS == 0.0f) totalRate = arrival / 60.0f;
    else totalRate = tTPS / (float)K;
    long long numReplicas = go_f64_to_int(ceil((double)totalRate / (double)rateStar));
    if (numReplicas < (long long)sys.srv_min_replicas[s]) numReplicas = sys.srv_min_replicas[s];
    const long long totalNumInstances = go_muli(num_instances(sys, m, a), numReplicas);
    const float cost = sys.acc_cost[a] * (float)totalNumInstances;
    const float rate = totalRate / (float)numReplicas;
    int st = qa.analyze(rate, metrics);
    steps += qa.steps;
    if (qa.fault) { fault = qa.fault; return false; }
    if (st != WVA_CAND_OK) return false;
    out.acc = a; out.numReplicas = numReplicas; out.batchSize = N;
    out.cost = cost; out.itl = metrics.avg_token_time;
    out.ttft = metrics.avg_wait_time + metrics.avg_prefill_time;
    out.rho = metrics.rho; out.maxArrv = rateStar / 1000.0f;
    out.value = cost;
    return true;
}

// Server.Calculate for one pair: candidate filter + CreateAllocation + transition penalty (server.go:55-67)
__device__ bool calculate_pair(const DevSystem& sys, int s, int a, double* scratch, AllocRec& out, int& fault,
                               unsigned long long& steps) {
    fault = 0;
    if (!is_candidate_accel(sys, s, a)) return false;
    if (!create_allocation(sys, s, a, scratch, out, fault, steps)) return false;
    out.value = transition_penalty(sys.srv_cur_acc[s], sys.srv_cur_replicas[s], sys.srv_cur_cost[s], out.acc,
                                   out.numReplicas, out.cost);
    return true;
}

// ---------------------------------------------------------------------------------------
// k_pairs: one thread per pair.  `order` (optional) permutes pair ids so that a warp holds pairs
// of similar chain length (sorted by N on the device beforehand) — lanes then finish together.
// Literal variant: `list` holds the pair ids that need the materialised path, scratch_off their
// p[] offsets.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_pairs(DevSystem sys, int s0, int nPairs, const int* __restrict__ order, DevAllocs out, unsigned char* feasible,
        int* slow_list, int* slow_count, unsigned long long* step_counter) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long steps = 0;
    if (t < nPairs) {
        int pid = order ? order[t] : t;           // pair id relative to shard: (s - s0) * A + a
        int s = s0 + pid / sys.A, a = pid % sys.A;
        AllocRec rec = empty_alloc();
        int fault = 0;
        bool ok = calculate_pair(sys, s, a, nullptr, rec, fault, steps);
        size_t gi = (size_t)s * sys.A + a;
        if (fault == 1) {
            int k = atomicAdd(slow_count, 1);
            slow_list[k] = pid;                   // capacity = nPairs
            ok = false;
        }
        if (!ok) rec = empty_alloc();
        store_alloc(out, gi, rec);
        feasible[gi] = ok ? 1 : 0;
    }
    // warp-reduce the step counter
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
    if ((threadIdx.x & 31) == 0 && steps) atomicAdd(step_counter, steps);
}

__global__ void __launch_bounds__(64)
k_pairs_literal(DevSystem sys, int s0, const int* __restrict__ list, int nList, double* scratch,
                const long long* __restrict__ scratch_off, DevAllocs out, unsigned char* feasible,
                unsigned long long* step_counter) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nList) return;
    int pid = list[t];
    int s = s0 + pid / sys.A, a = pid % sys.A;
    AllocRec rec = empty_alloc();
    int fault = 0;
    unsigned long long steps = 0;
    bool ok = calculate_pair(sys, s, a, scratch + scratch_off[t], rec, fault, steps);
    if (!ok) rec = empty_alloc();
    size_t gi = (size_t)s * sys.A + a;
    store_alloc(out, gi, rec);
    feasible[gi] = ok ? 1 : 0;
    atomicAdd(step_counter, steps);
}

// ---------------------------------------------------------------------------------------
// k_pairs_warp: one WARP per (server, accelerator) pair — the latency-oriented variant used when
// there are fewer pairs than the GPU has warps.  core.CreateAllocation is a chain of ~25-205
// dependent Solves (two bisections, then two Analyzes); a single thread runs them one after the
// other.  Here:
//   * the pair's {service rate, refined reciprocal} table is built once by the 32 lanes into global
//     memory (N x 16 B), so every chain step, ramp included, is DMUL + DMUL + 2 DFMA;
//   * the TTFT and ITL bisections of QueueAnalyzer.Size run concurrently on the two half-warps
//     (they only share the model through the stale-rho validity test, which is vacuous for
//     K = 11N >= 11 and lambda >= 0: SURVEY Appendix D.3);
//   * each bisection is evaluated speculatively: the interval endpoints are known float32 values, so
//     the midpoints of the next 4 levels of the bisection tree (15 nodes) are known too; 15 lanes
//     evaluate them at once and the warp then walks the tree with the reference's own comparisons.
//     The sequence of midpoints, iteration count (<= 100) and early exits are exactly those of
//     analyzer.BinarySearch (utils.go:26-70); evaluations off the realised path are discarded;
//   * the two trailing Analyzes (queueanalyzer.go:237-241, allocation.go:148) run side by side: the
//     second one's rate depends on the first one's throughput, which equals lambda*1000 exactly
//     whenever float32(p[K]) < 2^-25 — it is evaluated under that guess and re-done if wrong.
// Lanes call solve_uni together (idle lanes pass active = false), so the passes stay converged.
// ---------------------------------------------------------------------------------------

// x of node `node` (1-based heap index; children 2j = "xMax = x", 2j+1 = "xMin = x") of the
// bisection tree over [lo, hi]
__device__ __forceinline__ float bisect_node_x(float lo, float hi, int node) {
    const int depth = 31 - __clz(node);
    float x = 0.5f * (lo + hi);
    for (int l = depth - 1; l >= 0; --l) {
        if ((node >> l) & 1) lo = x; else hi = x;
        x = 0.5f * (lo + hi);
    }
    return x;
}

struct WarpPair {
    ServFormula sv; ProvTableF pv;
    double* pstore;                 // this lane's column of the warp's chain-value buffer, or nullptr
    bool cert;
    int N, K; long long inTok, outTok; bool tame;
    float rateMin, rateMax;
};

// one evaluation per lane (EvalTTFT / EvalITL / Analyze share the Solve); returns false when the
// lane's chain needs the materialised path.  certOnly: a speculative evaluation — if the certified tail
// cannot decide, report `uncertain` instead of running the long exact tail (the caller re-evaluates the
// node exactly only if the bisection actually walks through it).
__device__ __forceinline__ bool warp_solve(const WarpPair& wp, bool active, float x, SolveStats& st, unsigned long long& steps,
                                           bool& valid, bool certOnly, bool& uncertain) {
    int rc = solve_uni(wp.pv, wp.N, wp.K, x, wp.tame, st, steps, active, wp.pstore, wp.cert, certOnly);
    valid = true; uncertain = false;
    if (!active) return true;
    if (x < 0.0f) { valid = false; return true; }                       // queuemodel.go:31 (stale rho is in [0,1] < K)
    if (rc == WVA_SOLVE_UNCERTAIN) { uncertain = true; return true; }
    if (rc == WVA_SOLVE_CAREFUL) rc = solve_stream(wp.sv, (long long)wp.N, (long long)wp.K, x, wp.tame, st, steps);
    return rc == WVA_SOLVE_OK;
}
__device__ __forceinline__ float eval_y(const WarpPair& wp, int kind, const SolveStats& st) {
    float effConc = effective_concurrency(st.avgServTime, wp.sv.sp, wp.inTok, wp.outTok, wp.N);
    if (kind == 0) return st.avgWaitTime + prefill_time(wp.sv.sp, wp.inTok, effConc);   // EvalTTFT :270-279
    return decode_time(wp.sv.sp, effConc);                                              // EvalITL  :283-290
}
// QueueAnalyzer.Analyze's metrics from a Solve (queueanalyzer.go:152-173)
__device__ __forceinline__ void metrics_from(const WarpPair& wp, const SolveStats& st, wva_metrics& m) {
    float effConc = effective_concurrency(st.avgServTime, wp.sv.sp, wp.inTok, wp.outTok, wp.N);
    float rho = st.avgNumInServers / (float)wp.N;
    rho = go_minf(go_maxf(rho, 0.0f), 1.0f);
    m.throughput = st.throughput * 1000.0f;
    m.avg_resp_time = st.avgRespTime;
    m.avg_wait_time = st.avgWaitTime;
    m.avg_num_in_serv = st.avgNumInServers;
    m.avg_prefill_time = prefill_time(wp.sv.sp, wp.inTok, effConc);
    m.avg_token_time = decode_time(wp.sv.sp, effConc);
    m.max_rate = wp.rateMax;
    m.rho = rho;
}
__device__ __forceinline__ SolveStats shfl_stats(const SolveStats& s, int src) {
    SolveStats o;
    o.rho = __shfl_sync(0xffffffffu, s.rho, src);
    o.avgNumInServers = __shfl_sync(0xffffffffu, s.avgNumInServers, src);
    o.avgNumInSystem = __shfl_sync(0xffffffffu, s.avgNumInSystem, src);
    o.throughput = __shfl_sync(0xffffffffu, s.throughput, src);
    o.avgRespTime = __shfl_sync(0xffffffffu, s.avgRespTime, src);
    o.avgServTime = __shfl_sync(0xffffffffu, s.avgServTime, src);
    o.avgWaitTime = __shfl_sync(0xffffffffu, s.avgWaitTime, src);
    return o;
}

#define WVA_PAIRS_WARP_THREADS 128
__global__ void __launch_bounds__(WVA_PAIRS_WARP_THREADS)
k_pairs_warp(DevSystem sys, int s0, int nPairs, const long long* __restrict__ tabOff, double2* tabs, DevAllocs out,
             unsigned char* feasible, int* slow_list, int* slow_count, unsigned long long* step_counter,
             int smemEntriesPerWarp, double* pbuf, long long pbufStrideK, unsigned long long* dbg) {
    extern __shared__ __align__(16) unsigned char pairs_smem[];
    const long long tStart = clock64();
    int activeRounds = 0, totalRounds = 0;
    const int pid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // warp-uniform
    const int lane = threadIdx.x & 31;
    if (pid >= nPairs) return;
    const int s = s0 + pid / sys.A, a = pid % sys.A;
    const size_t gi = (size_t)s * sys.A + a;
    unsigned long long steps = 0;
    AllocRec rec = empty_alloc();
    bool ok = false;          // CreateAllocation != nil
    bool toSlow = false;      // needs the materialised path (whole pair re-done by k_pairs_literal)

    // every lane runs the scalar part redundantly (uniform control flow)
    do {
        if (!is_candidate_accel(sys, s, a) || !pair_lookups_ok(sys, s, a)) break;
        const float arrival = sys.srv_arrival_rpm[s];
        const long long inTok = sys.srv_in_tokens[s], outTok = sys.srv_out_tokens[s];
        if (arrival == 0.0f || outTok == 0) { rec = zero_load_allocation(sys, s, a); ok = true; break; }
        const int m = sys.srv_model[s];
        const size_t pi = (size_t)m * sys.A + a;
        const long long Kt = outTok;
        long long N;
        if (sys.srv_max_batch[s] > 0) N = sys.srv_max_batch[s];
        else { N = go_divi(go_muli(sys.perf_max_batch[pi], sys.perf_at_tokens[pi]), Kt); if (N < 1) N = 1; }
        const long long maxQueue = go_muli(N, WVA_MAX_QUEUE_TO_BATCH_RATIO);
        if (!config_ok(N, maxQueue, inTok, Kt)) break;
        if (tabOff[pid] < 0 || N > (1LL << 26)) { toSlow = true; break; }   // no table budget: thread kernel handles it
        WarpPair wp;
        ServiceParms sp; sp.alpha = sys.perf_alpha[pi]; sp.beta = sys.perf_beta[pi];
        sp.gamma = sys.perf_gamma[pi]; sp.delta = sys.perf_delta[pi];
        wp.sv.init(sp, inTok, Kt);
        wp.N = (int)N; wp.K = (int)(maxQueue + N); wp.inTok = inTok; wp.outTok = Kt;
        wp.tame = tame_parms(sp, inTok, Kt);
        wp.cert = sys.cert != 0;
        // the table lives in shared memory when it fits (29-cycle loads in the ramp), else in HBM
        double2* tab = (wp.N <= smemEntriesPerWarp)
                           ? reinterpret_cast<double2*>(pairs_smem) + (size_t)(threadIdx.x >> 5) * smemEntriesPerWarp
                           : tabs + tabOff[pid];
        wp.pstore = (pbuf && (long long)wp.K + 1 <= pbufStrideK) ? pbuf + (size_t)pid * (size_t)pbufStrideK * 32 + lane : nullptr;
        // BuildModel (queueanalyzer.go:99-131): the table, cooperatively
        bool bad = false;
        for (int i = lane; i < wp.N; i += 32) {
            float r = wp.sv.rate(i + 1);
            if (!(r > 0.0f) || !(r < CUDART_INF_F)) bad = true;
            double d = (double)r;
            tab[i] = make_double2(d, rcp_refined(d));
        }
        __syncwarp();
        if (__any_sync(0xffffffffu, bad)) { toSlow = true; break; }
        wp.pv.tab = tab; wp.pv.sf = &wp.sv;
        {
            float lambdaMin = wp.sv.rate(1) * WVA_EPSILON;
            float lambdaMax = wp.sv.rate(wp.N) * (1.0f - WVA_EPSILON);
            wp.rateMin = lambdaMin * 1000.0f; wp.rateMax = lambdaMax * 1000.0f;
        }
        const float tTTFT = sys.srv_slo_ttft[s], tITL = sys.srv_slo_itl[s], tTPS = sys.srv_slo_tps[s];
        // ---- QueueAnalyzer.Size (queueanalyzer.go:185-255) ------------------------------------
        if (tITL < 0.0f || tTTFT < 0.0f || tTPS < 0.0f) break;
        const float lambdaMin = wp.rateMin / 1000.0f, lambdaMax = wp.rateMax / 1000.0f;
        const int half = lane >> 4, hl = lane & 15, base = lane & 16;
        const float target = half == 0 ? tTTFT : tITL;
        const bool searching = target > 0.0f;          // this half-warp's bisection is requested
        float lo = lambdaMin, hi = lambdaMax;
        float xStar = lambdaMax;                       // lambdaStar when the target is disabled
        int ind = 0, iters = 0;
        bool done = !searching, failed = false, inc = false;
        if (searching && lambdaMin > lambdaMax) { failed = true; done = true; }     // utils.go:29-31
        SolveStats st; st.rho = st.avgNumInServers = st.avgNumInSystem = st.throughput = st.avgRespTime = st.avgServTime = st.avgWaitTime = 0.0f;
        float lastX = -1.0f;                           // this lane's most recent evaluation point
        float ylo = 0.0f, yhi = 0.0f;                  // eval(lo), eval(hi): always known after the first round,
                                                       // so a midpoint that rounds onto an endpoint costs nothing
        bool first = true;
        int rounds = 0;
        while (__any_sync(0xffffffffu, !done)) {
            ++rounds;
            // --- choose this lane's evaluation point ---
            int node = 0; bool act = false; float x = 0.0f;
            if (!done) {
                if (first) {
                    if (hl == 0) { act = true; x = lo; }
                    else if (hl == 1) { act = true; x = hi; }
                    else if (hl <= 8) { node = hl - 1; act = true; x = bisect_node_x(lo, hi, node); }
                } else if (hl >= 1) { node = hl; act = true; x = bisect_node_x(lo, hi, node); }
                // never evaluate beyond the reference's iteration budget
                if (node > 0 && (31 - __clz(node)) >= WVA_BISECT_MAXIT - iters) act = false;
                // a midpoint equal to an endpoint of its own interval has a known value.  Endpoints of
                // deeper nodes are midpoints of shallower ones, so it is enough to test against the
                // node's own [lo', hi']: recompute them along the path.
                if (node > 0 && act) {
                    float l2 = lo, h2 = hi, xm = 0.5f * (lo + hi);
                    for (int l = (31 - __clz(node)) - 1; l >= 0; --l) { if ((node >> l) & 1) l2 = xm; else h2 = xm; xm = 0.5f * (l2 + h2); }
                    if (xm == l2 || xm == h2) act = false;      // resolved from the memo while walking
                }
            }
            bool valid, unc;
            if (__any_sync(0xffffffffu, act)) ++activeRounds;
            ++totalRounds;
            // tree nodes are speculative (certOnly); the two boundary evaluations are needed for sure
            bool solved = warp_solve(wp, act, x, st, steps, valid, node > 0, unc);
            if (act && !unc) lastX = x;
            if (__any_sync(0xffffffffu, act && !solved)) { toSlow = true; break; }
            float y = act && valid && !unc ? eval_y(wp, half, st) : 0.0f;
            // --- walk the realised path (uniform inside each half-warp) ---
            if (first) {
                const float yb0 = __shfl_sync(0xffffffffu, y, base + 0), yb1 = __shfl_sync(0xffffffffu, y, base + 1);
                const bool v0 = __shfl_sync(0xffffffffu, (int)valid, base + 0), v1 = __shfl_sync(0xffffffffu, (int)valid, base + 1);
                if (!done) {
                    if (!v0) { failed = true; done = true; }
                    else if (within_tolerance(yb0, target, WVA_BISECT_TOL)) { xStar = lo; ind = 0; done = true; }
                    else if (!v1) { failed = true; done = true; }
                    else if (within_tolerance(yb1, target, WVA_BISECT_TOL)) { xStar = hi; ind = 0; done = true; }
                    else {
                        inc = yb0 < yb1;
                        ylo = yb0; yhi = yb1;
                        if ((inc && target < yb0) || (!inc && target > yb0)) { xStar = lo; ind = -1; done = true; }
                        else if ((inc && target > yb1) || (!inc && target < yb1)) { xStar = hi; ind = +1; done = true; }
                    }
                }
            }
            {
                const int levels = first ? 3 : 4;
                int cur = 1, lvl = 0;
                bool slowExit = false;
                for (;;) {
                    bool pend = false;                 // this half-warp walked into a node whose speculative evaluation was uncertain
                    for (int l = 0; l < levels; ++l) {
                        const int src = base + (first ? cur + 1 : cur);
                        float ys = __shfl_sync(0xffffffffu, y, src);
                        bool vs = __shfl_sync(0xffffffffu, (int)valid, src);
                        const bool us = __shfl_sync(0xffffffffu, (int)unc, src);
                        if (!done && !pend && l == lvl) {
                            if (iters == WVA_BISECT_MAXIT) { done = true; }
                            else {
                                const float xs = 0.5f * (lo + hi);
                                bool known = true;
                                if (xs == lo) { ys = ylo; vs = true; } else if (xs == hi) { ys = yhi; vs = true; }
                                else if (us) known = false;
                                if (!known) pend = true;           // needs the exact chain: evaluated below, then the walk resumes here
                                else {
                                    ++iters; ++lvl;
                                    if (!vs) { failed = true; done = true; }
                                    else {
                                        xStar = xs;
                                        if (within_tolerance(ys, target, WVA_BISECT_TOL)) done = true;
                                        else if ((inc && target < ys) || (!inc && target > ys)) { hi = xs; yhi = ys; cur = 2 * cur; }
                                        else { lo = xs; ylo = ys; cur = 2 * cur + 1; }
                                        if (!done && iters == WVA_BISECT_MAXIT) done = true;
                                    }
                                }
                            }
                        }
                    }
                    if (!__any_sync(0xffffffffu, pend)) break;
                    // exact evaluation of the pending node(s): the lane that owns node `cur` of a pending half-warp
                    const bool mine = pend && (hl == (first ? cur + 1 : cur));
                    bool v2, u2;
                    ++activeRounds;
                    bool solved2 = warp_solve(wp, mine, x, st, steps, v2, false, u2);
                    if (__any_sync(0xffffffffu, mine && !solved2)) { slowExit = true; break; }
                    if (mine) { valid = v2; unc = false; lastX = x; y = v2 ? eval_y(wp, half, st) : 0.0f; }
                }
                if (slowExit) { toSlow = true; break; }
            }
            first = false;
        }
        if (lane == 0) { atomicAdd(step_counter + 1, (unsigned long long)rounds); atomicMax(step_counter + 2, (unsigned long long)rounds); }
        if (toSlow) break;
        // results of the two bisections
        const bool failT = __shfl_sync(0xffffffffu, (int)failed, 0), failI = __shfl_sync(0xffffffffu, (int)failed, 16);
        const int indT = __shfl_sync(0xffffffffu, ind, 0), indI = __shfl_sync(0xffffffffu, ind, 16);
        const float lTTFT = __shfl_sync(0xffffffffu, xStar, 0), lITL = __shfl_sync(0xffffffffu, xStar, 16);
        if ((tTTFT > 0.0f && (failT || indT < 0))) break;                  // :205-214
        if ((tITL > 0.0f && (failI || indI < 0))) break;                   // :218-228
        float lTPS = lambdaMax;
        if (tTPS > 0.0f) lTPS = lambdaMax * (1.0f - WVA_STABILITY_SAFETY);
        const float lambda = go_minf(go_minf(lTTFT, lITL), lTPS);
        const float requestRate = lambda * 1000.0f;
        // ---- Analyze(requestRate) (:237-241) and Analyze(totalRate/replicas) (allocation.go:148) ----
        if (requestRate <= 0.0f || requestRate > wp.rateMax) break;        // :135-143 -> Size fails -> nil
        const float x1 = requestRate / 1000.0f;
        float totalRate;
        if (tTPS == 0.0f) totalRate = arrival / 60.0f;
        else totalRate = tTPS / (float)Kt;
        const int minRep = sys.srv_min_replicas[s];
        // guess: throughput == x1 (blocking probability below float32 resolution)
        const float rateStarGuess = x1 * 1000.0f;
        long long repGuess = go_f64_to_int(ceil((double)totalRate / (double)rateStarGuess));
        if (repGuess < (long long)minRep) repGuess = minRep;
        const float rate2Guess = totalRate / (float)repGuess;
        const bool rate2GuessOk = !(rate2Guess <= 0.0f) && !(rate2Guess > wp.rateMax);
        // reuse an evaluation already made at exactly x1 if some lane has one
        const unsigned have = __ballot_sync(0xffffffffu, lastX == x1);
        SolveStats st1, st2;
        bool valid1 = true, valid2 = true, solved = true;
        {
            const bool needX1 = (have == 0u);
            const bool actA = (lane == 0) && needX1, actB = (lane == 1) && rate2GuessOk;
            SolveStats stl = st;
            bool v;
            bool uu;
            solved = warp_solve(wp, actA || actB, lane == 0 ? x1 : rate2Guess / 1000.0f, stl, steps, v, false, uu);
            if (__any_sync(0xffffffffu, (actA || actB) && !solved)) { toSlow = true; break; }
            if (needX1) { st1 = shfl_stats(stl, 0); valid1 = __shfl_sync(0xffffffffu, (int)v, 0); }
            else { st1 = shfl_stats(st, __ffs(have) - 1); valid1 = true; }
            st2 = shfl_stats(stl, 1); valid2 = __shfl_sync(0xffffffffu, (int)v, 1);
        }
        if (!valid1) break;
        wva_metrics m1; metrics_from(wp, st1, m1);
        const float rateStar = m1.throughput;
        long long numReplicas = go_f64_to_int(ceil((double)totalRate / (double)rateStar));
        if (numReplicas < (long long)minRep) numReplicas = minRep;
        const long long totalNumInstances = go_muli(num_instances(sys, m, a), numReplicas);
        const float cost = sys.acc_cost[a] * (float)totalNumInstances;
        const float rate2 = totalRate / (float)numReplicas;
        if (rate2 <= 0.0f || rate2 > wp.rateMax) break;                    // Analyze error -> nil
        if (lane == 0) atomicAdd(step_counter + 3, (unsigned long long)((have == 0u ? 1 : 0) + ((rate2GuessOk && rate2 == rate2Guess) ? 0 : 100)));
        if (!(rate2GuessOk && rate2 == rate2Guess)) {
            // the guess missed: evaluate the real second rate
            SolveStats stl = st; bool v;
            bool uu;
            solved = warp_solve(wp, lane == 0, rate2 / 1000.0f, stl, steps, v, false, uu);
            if (__any_sync(0xffffffffu, lane == 0 && !solved)) { toSlow = true; break; }
            st2 = shfl_stats(stl, 0); valid2 = __shfl_sync(0xffffffffu, (int)v, 0);
        }
        if (!valid2) break;
        wva_metrics m2; metrics_from(wp, st2, m2);
        rec.acc = a; rec.numReplicas = numReplicas; rec.batchSize = N;
        rec.cost = cost; rec.itl = m2.avg_token_time;
        rec.ttft = m2.avg_wait_time + m2.avg_prefill_time;
        rec.rho = m2.rho; rec.maxArrv = rateStar / 1000.0f;
        rec.value = cost;
        ok = true;
    } while (false);

    if (ok) rec.value = transition_penalty(sys.srv_cur_acc[s], sys.srv_cur_replicas[s], sys.srv_cur_cost[s], rec.acc,
                                           rec.numReplicas, rec.cost);
    if (dbg && lane == 0) { dbg[2 * pid] = (unsigned long long)(clock64() - tStart); dbg[2 * pid + 1] = ((unsigned long long)totalRounds << 32) | (unsigned)activeRounds; }
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
    if (lane == 0) {
        if (toSlow) { ok = false; slow_list[atomicAdd(slow_count, 1)] = pid; }
        if (!ok) rec = empty_alloc();
        store_alloc(out, gi, rec);
        feasible[gi] = ok ? 1 : 0;
        if (steps) atomicAdd(step_counter, steps);
    }
}

// N per pair (0 = no queueing work) — used to order pairs by chain length and to size scratch.
__global__ void k_pair_batch(DevSystem sys, int s0, int nPairs, long long* __restrict__ nOut) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nPairs) return;
    nOut[t] = pair_batch_size(sys, s0 + t / sys.A, t % sys.A);
}
// bucket = position of the highest set bit of N (0 for N == 0): 0..63
__global__ void k_pair_bucket_hist(const long long* __restrict__ nIn, int nPairs, int* __restrict__ hist) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nPairs) return;
    long long n = nIn[t];
    int b = n <= 0 ? 0 : (64 - __clzll(n));
    atomicAdd(&hist[b], 1);
}
// scatter pair ids into descending-bucket order (heaviest first); order within a bucket is free
__global__ void k_pair_bucket_scatter(const long long* __restrict__ nIn, int nPairs, int* __restrict__ cursor,
                                      int* __restrict__ order) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nPairs) return;
    long long n = nIn[t];
    int b = n <= 0 ? 0 : (64 - __clzll(n));
    int pos = atomicAdd(&cursor[b], 1);
    order[pos] = t;
}

// ---------------------------------------------------------------------------------------
// Candidate sweep
// ---------------------------------------------------------------------------------------

// best candidate of a block / a list thread: key + the metrics a wva_grid_best needs
struct GridSlot { unsigned long long key; float cost, itl, ttft, rho; int sl; int pad; };

struct GridParams {
    int r_max, b_max;
    int r_chunk, n_rchunks;          // replicas per block / blocks per pair
    int s0, ns;                      // shard
    int pair_base;                   // (server, accelerator) pairs of the shard that precede the slice being swept
    wva_metrics* cube;               // [ns*A*r_max*b_max] or nullptr
    unsigned char* status;           // same extent or nullptr
    unsigned long long* keys;        // [ns] per-server argmin key (initialised to ~0)
    unsigned long long* counters;    // [0] steps executed, [1] algorithmic steps, [2] candidates analysed ok
    unsigned long long* slow_list; int* slow_count; int slow_cap;   // candidate ids (relative to shard) needing the literal path
    double2* pair_tab;               // [ns*A][b_max] {rate, refined reciprocal}: written by k_grid, read by the list kernel
    struct GridSlot* block_slot;     // [blocks of k_grid] best candidate of each block with its metrics
    struct GridSlot* list_slot;      // [heavy_cap + slow_cap] one per list-kernel thread
    int tail_cap;                    // tail steps a candidate may take inside k_grid before it is deferred
    unsigned long long* heavy_list; float* heavy_cost; int* heavy_count; int heavy_cap;   // deferred (long) chains
};

// order-preserving map float -> uint32 (ascending), -0 canonicalised by the caller
__device__ __forceinline__ unsigned sortable_f32(float v) {
    unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unsortable_f32(unsigned k) {
    unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}
// key = (value, accelerator, replicas, batch) lexicographic
__device__ __forceinline__ unsigned long long make_key(float value, int a, int r, int b) {
    return ((unsigned long long)sortable_f32(value) << 32) | ((unsigned long long)a << 24) |
           ((unsigned long long)(r - 1) << 14) | (unsigned long long)(b - 1);
}
#define WVA_KEY_NONE 0xffffffffffffffffULL
#define WVA_GRID_MAX_A 256
#define WVA_GRID_MAX_R 1024
#define WVA_GRID_MAX_B 8192

// Per-server constants of the sweep, computed once per block
struct GridServer {
    float totalRate;                 // allocation.go:134-139
    float sloTTFT, sloITL, sloTPS;
    long long inTok, outTok;
    int minReplicas;
    int curAcc, curRep; float curCost;
    float accCost; long long numInst;
    ServiceParms sp;
    bool cert;
};

// QueueAnalyzer.Analyze (queueanalyzer.go:134-174) on a fresh analyzer with MaxBatchSize b,
// MaxQueueSize 10 b, service rates from the shared-memory table.  A fresh model has p[0] = 0 so
// the validity test (queuemodel.go:30-31) sees rho = 1: valid iff 1 < K and lambda >= 0.
__device__ __forceinline__ int analyze_table(const ServTable& tb, const GridServer& gs, int b, float rate, bool tame,
                                             int tailCap, wva_metrics& m, float& rateTPS, unsigned long long& steps,
                                             float& deferCost) {
    const int K = b * WVA_MAX_QUEUE_TO_BATCH_RATIO + b;
    const float lambdaMax = tb.rateF[b - 1] * (1.0f - WVA_EPSILON);
    const float rateMax = lambdaMax * 1000.0f;
    const float lamMaxBack = rateMax / 1000.0f;
    rateTPS = (lamMaxBack * (1.0f - WVA_STABILITY_SAFETY)) * 1000.0f;    // TargetRate.RateTargetTPS, :231-234,:246
    if (rate <= 0.0f) return WVA_CAND_ERR_RATE_LE0;
    if (rate > rateMax) return WVA_CAND_ERR_RATE_MAX;
    const float lambda = rate / 1000.0f;
    if ((1.0f >= (float)K) || (lambda < 0.0f)) return WVA_CAND_ERR_MODEL;
    SolveStats st;
    ProvTable pv; pv.rateD = tb.rateD; pv.rcp = tb.rcp; pv.rateF_ = tb.rateF;
    int rc = solve_fast(pv, b, K, lambda, tame, tailCap, st, steps, deferCost, gs.cert);
    if (rc == WVA_SOLVE_DEFER) return -2;                                 // long chain: heavy kernel
    if (rc == WVA_SOLVE_CAREFUL) rc = solve_stream_table(tb, b, K, lambda, tame, st, steps);
    if (rc != WVA_SOLVE_OK) return -1;                                    // literal path
    float effConc = effective_concurrency(st.avgServTime, gs.sp, gs.inTok, gs.outTok, b);
    float rho = st.avgNumInServers / (float)b;
    rho = go_minf(go_maxf(rho, 0.0f), 1.0f);
    m.throughput = st.throughput * 1000.0f;
    m.avg_resp_time = st.avgRespTime;
    m.avg_wait_time = st.avgWaitTime;
    m.avg_num_in_serv = st.avgNumInServers;
    m.avg_prefill_time = prefill_time(gs.sp, gs.inTok, effConc);
    m.avg_token_time = decode_time(gs.sp, effConc);
    m.max_rate = rateMax;
    m.rho = rho;
    return WVA_CAND_OK;
}

// SLO / replica constraints of a candidate and its key (WVA_KEY_NONE when infeasible)
__device__ __forceinline__ unsigned long long candidate_key(const GridServer& gs, int a, int r, int b, float rate,
                                                            float rateTPS, const wva_metrics& m, bool& feasible) {
    const float ttft = m.avg_wait_time + m.avg_prefill_time;
    const float itl = m.avg_token_time;
    feasible = (!(gs.sloTTFT > 0.0f) || ttft <= gs.sloTTFT) && (!(gs.sloITL > 0.0f) || itl <= gs.sloITL) &&
               (!(gs.sloTPS > 0.0f) || rate <= rateTPS) && (r >= gs.minReplicas);
    if (!feasible) return WVA_KEY_NONE;
    const float cost = gs.accCost * (float)go_muli(gs.numInst, (long long)r);
    float value = transition_penalty(gs.curAcc, gs.curRep, gs.curCost, a, (long long)r, cost);
    value = value + 0.0f;
    if (value != value) return WVA_KEY_NONE;
    return make_key(value, a, r, b);
}

__device__ __forceinline__ void load_grid_server(const DevSystem& sys, int s, int a, GridServer& gs) {
    const int m = sys.srv_model[s];
    const size_t pi = (size_t)m * sys.A + a;
    gs.sp.alpha = sys.perf_alpha[pi]; gs.sp.beta = sys.perf_beta[pi];
    gs.sp.gamma = sys.perf_gamma[pi]; gs.sp.delta = sys.perf_delta[pi];
    gs.inTok = sys.srv_in_tokens[s]; gs.outTok = sys.srv_out_tokens[s];
    gs.sloTTFT = sys.srv_slo_ttft[s]; gs.sloITL = sys.srv_slo_itl[s]; gs.sloTPS = sys.srv_slo_tps[s];
    gs.totalRate = (gs.sloTPS == 0.0f) ? sys.srv_arrival_rpm[s] / 60.0f : gs.sloTPS / (float)gs.outTok;
    gs.minReplicas = sys.srv_min_replicas[s];
    gs.curAcc = sys.srv_cur_acc[s]; gs.curRep = sys.srv_cur_replicas[s]; gs.curCost = sys.srv_cur_cost[s];
    gs.accCost = sys.acc_cost[a];
    gs.numInst = num_instances(sys, m, a);
    gs.cert = sys.cert != 0;
}

// Block = one (server, accelerator) pair x one chunk of replica counts.  The pair's service-rate
// table (float, double and refined reciprocal, 20 B per batch size) is built once into shared
// memory; warps then pull (replicas, 32 consecutive batch sizes) work items from a shared counter:
// lanes of a warp share lambda and differ only by batch size, so their trip counts are close.
// Chains whose tail outlasts gp.tail_cap are deferred to k_grid_list.
#define WVA_GRID_THREADS 256
__global__ void __launch_bounds__(WVA_GRID_THREADS, 3)
k_grid(DevSystem sys, GridParams gp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* rateD = reinterpret_cast<double*>(smem_raw);
    double* rcp = rateD + gp.b_max;
    float* rateF = reinterpret_cast<float*>(rcp + gp.b_max);
    __shared__ int sh_item, sh_nGood;
    __shared__ unsigned long long sh_key;
    __shared__ unsigned long long sh_cnt[3];

    const int pairSlice = blockIdx.x / gp.n_rchunks;       // pair index inside the slice
    const int pairLocal = gp.pair_base + pairSlice;        // (s - s0) * A + a
    const int rchunk = blockIdx.x % gp.n_rchunks;
    const int sl = pairLocal / sys.A, a = pairLocal % sys.A;
    const int s = gp.s0 + sl;
    const int r_lo = rchunk * gp.r_chunk + 1;
    const int r_hi = min(gp.r_max, r_lo + gp.r_chunk - 1);
    const int B = gp.b_max;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        sh_item = 0; sh_nGood = B; sh_key = WVA_KEY_NONE; sh_cnt[0] = sh_cnt[1] = sh_cnt[2] = 0;
        gp.block_slot[blockIdx.x].key = WVA_KEY_NONE;
    }

    const bool pairOk = pair_lookups_ok(sys, s, a) && is_candidate_accel(sys, s, a);
    GridServer gs;
    int blockStatus = WVA_CAND_OK;
    if (!pairOk) blockStatus = WVA_CAND_ERR_PAIR;
    else {
        load_grid_server(sys, s, a, gs);
        if (gs.inTok < 0 || gs.outTok < 1 || gs.sloTTFT < 0.0f || gs.sloITL < 0.0f || gs.sloTPS < 0.0f)
            blockStatus = WVA_CAND_ERR_CONFIG;
    }
    const size_t candBase = ((size_t)pairLocal * gp.r_max) * (size_t)B;
    if (blockStatus != WVA_CAND_OK) {
        // whole block shares one error status; zero metrics
        if (gp.status || gp.cube) {
            const size_t n = (size_t)(r_hi - r_lo + 1) * B;
            const size_t off = candBase + (size_t)(r_lo - 1) * B;
            for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
                if (gp.status) gp.status[off + i] = (unsigned char)blockStatus;
                if (gp.cube) { float4 z = make_float4(0, 0, 0, 0); float4* c = reinterpret_cast<float4*>(&gp.cube[off + i]); c[0] = z; c[1] = z; }
            }
        }
        return;
    }
    __syncthreads();

    // ---- stage the service-rate table (and publish it for the deferred-chain kernel) ------------
    ServFormula sf; sf.init(gs.sp, gs.inTok, gs.outTok);
    double2* gtab = (rchunk == 0) ? gp.pair_tab + (size_t)pairSlice * B : nullptr;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        float r = sf.rate(i + 1);
        rateF[i] = r;
        double d = (double)r;
        double y = rcp_refined(d);
        rateD[i] = d;
        rcp[i] = y;
        if (gtab) gtab[i] = make_double2(d, y);
        if (!(r > 0.0f) || !(r < CUDART_INF_F)) atomicMin(&sh_nGood, i);
    }
    const bool tame = tame_parms(gs.sp, gs.inTok, gs.outTok);
    __syncthreads();
    ServTable tb; tb.rateF = rateF; tb.rateD = rateD; tb.rcp = rcp;
    const int nGood = sh_nGood;

    // ---- sweep ------------------------------------------------------------------------------
    const int bChunks = (B + 31) / 32;
    const int nItems = (r_hi - r_lo + 1) * bChunks;
    unsigned long long bestKey = WVA_KEY_NONE;
    float bestItl = 0.0f, bestTtft = 0.0f, bestRho = 0.0f;
    unsigned long long steps = 0, algSteps = 0, okCount = 0;
    for (;;) {
        int item;
        if (lane == 0) item = atomicAdd(&sh_item, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nItems) break;
        // heaviest items first: large batch sizes have the longest ramps
        const int bc = bChunks - 1 - (item % bChunks);
        const int r = r_lo + item / bChunks;
        const int b = bc * 32 + lane + 1;
        if (b > B) continue;
        const size_t ci = candBase + (size_t)(r - 1) * B + (size_t)(b - 1);
        const float rate = gs.totalRate / (float)r;
        wva_metrics m;
        m.throughput = m.avg_resp_time = m.avg_wait_time = m.avg_num_in_serv = 0.0f;
        m.avg_prefill_time = m.avg_token_time = m.max_rate = m.rho = 0.0f;
        float rateTPS = 0.0f;
        int st;
        bool feasible = false;
        float deferCost = 0.0f;
        bool deferred = false;
        if (b > nGood) st = -1;
        else {
            // one call site: the second trip (cap 0) only happens when the deferred list is full
            // short chains (K <= 1024) never leave this kernel: deferral only pays for long tails
            for (int cap = (11 * b <= 1024) ? 0 : gp.tail_cap;; cap = 0) {
                st = analyze_table(tb, gs, b, rate, tame, cap, m, rateTPS, steps, deferCost);
                if (st != -2) break;
                int k = atomicAdd(gp.heavy_count, 1);
                if (k < gp.heavy_cap) { gp.heavy_list[k] = (unsigned long long)ci; gp.heavy_cost[k] = deferCost; deferred = true; break; }
            }
        }
        if (deferred) continue;      // long chain: k_grid_list groups chains of similar length
        if (st == -1) {
            // literal path: queue the candidate, outputs come from k_grid_list
            int k = atomicAdd(gp.slow_count, 1);
            if (k < gp.slow_cap) gp.slow_list[k] = (unsigned long long)ci;
            continue;
        }
        if (st == WVA_CAND_OK) {
            okCount++;
            algSteps += 2ULL * (unsigned long long)(11 * b + 1);
            unsigned long long key = candidate_key(gs, a, r, b, rate, rateTPS, m, feasible);
            if (key < bestKey) {
                bestKey = key; bestItl = m.avg_token_time; bestTtft = m.avg_wait_time + m.avg_prefill_time; bestRho = m.rho;
            }
        } else {
            m.throughput = m.avg_resp_time = m.avg_wait_time = m.avg_num_in_serv = 0.0f;
            m.avg_prefill_time = m.avg_token_time = m.max_rate = m.rho = 0.0f;
        }
        if (gp.cube) {
            float4* c = reinterpret_cast<float4*>(&gp.cube[ci]);
            c[0] = make_float4(m.throughput, m.avg_resp_time, m.avg_wait_time, m.avg_num_in_serv);
            c[1] = make_float4(m.avg_prefill_time, m.avg_token_time, m.max_rate, m.rho);
        }
        if (gp.status) gp.status[ci] = (unsigned char)(st | (feasible ? WVA_CAND_FEASIBLE : 0));
    }

    // ---- warp-shuffle then block argmin; the owner of the block minimum publishes its metrics ----
    unsigned long long warpKey = bestKey;
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, warpKey, o);
        if (other < warpKey) warpKey = other;
        steps += __shfl_down_sync(0xffffffffu, steps, o);
        algSteps += __shfl_down_sync(0xffffffffu, algSteps, o);
        okCount += __shfl_down_sync(0xffffffffu, okCount, o);
    }
    if (lane == 0) {
        if (warpKey != WVA_KEY_NONE) atomicMin(&sh_key, warpKey);
        atomicAdd(&sh_cnt[0], steps); atomicAdd(&sh_cnt[1], algSteps); atomicAdd(&sh_cnt[2], okCount);
    }
    __syncthreads();
    const unsigned long long blockKey = sh_key;
    if (blockKey != WVA_KEY_NONE && bestKey == blockKey) {          // keys are unique per candidate: one owner
        GridSlot sl_; sl_.key = blockKey; sl_.itl = bestItl; sl_.ttft = bestTtft; sl_.rho = bestRho; sl_.sl = sl; sl_.pad = 0;
        const int r = (int)((blockKey >> 14) & 0x3ff) + 1;
        sl_.cost = gs.accCost * (float)go_muli(gs.numInst, (long long)r);
        gp.block_slot[blockIdx.x] = sl_;
        atomicMin(&gp.keys[sl], blockKey);
    }
    if (threadIdx.x == 0) {
        atomicAdd(&gp.counters[0], sh_cnt[0]); atomicAdd(&gp.counters[1], sh_cnt[1]); atomicAdd(&gp.counters[2], sh_cnt[2]);
    }
}

// ---------------------------------------------------------------------------------------
// k_grid_rows: the sweep with ONE THREAD PER ROW (server, accelerator, replicas).
//
// All candidates of a row share lambda = totalRate / r, and the chain of batch size b uses
// servRate[min(n, b-1)]: its first b steps are the first b steps of the chain of every larger batch
// size.  So a row needs ONE ramp p[1..B]; candidate b is evaluated at the moment the ramp reaches state b,
// from (p[b], sum_{i<=b} p[i], sum_{i<=b} i p[i]) and the certified closed-form tail (cert_eval) — O(1) per
// candidate instead of O(b) + tail.  When the ramp itself dies out (p[n] below 2^-68 of every aggregate
// with all later ratios <= 0.998), the remaining candidates of the row share the frozen sums.
// A candidate whose certificate fails (ambiguous float32 rounding, ill-conditioned closed form, value
// window left) is appended to the deferred list and evaluated by the exact chain in k_grid_list.
// Lanes of a warp hold consecutive r of the same pair: they step n together, so the shared-memory
// table reads are broadcasts.
// ---------------------------------------------------------------------------------------
#define WVA_ROWS_THREADS 64
__global__ void __launch_bounds__(WVA_ROWS_THREADS)
k_grid_rows(DevSystem sys, GridParams gp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* rateD = reinterpret_cast<double*>(smem_raw);
    double* rcp = rateD + gp.b_max;
    float* rateF = reinterpret_cast<float*>(rcp + gp.b_max);
    __shared__ int sh_nGood;
    __shared__ unsigned long long sh_key;
    __shared__ unsigned long long sh_cnt[3];

    const int pairSlice = blockIdx.x;
    const int pairLocal = gp.pair_base + pairSlice;
    const int sl = pairLocal / sys.A, a = pairLocal % sys.A;
    const int s = gp.s0 + sl;
    const int B = gp.b_max, R = gp.r_max;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        sh_nGood = B; sh_key = WVA_KEY_NONE; sh_cnt[0] = sh_cnt[1] = sh_cnt[2] = 0;
        gp.block_slot[blockIdx.x].key = WVA_KEY_NONE;
    }
    const bool pairOk = pair_lookups_ok(sys, s, a) && is_candidate_accel(sys, s, a);
    GridServer gs;
    int blockStatus = WVA_CAND_OK;
    if (!pairOk) blockStatus = WVA_CAND_ERR_PAIR;
    else {
        load_grid_server(sys, s, a, gs);
        if (gs.inTok < 0 || gs.outTok < 1 || gs.sloTTFT < 0.0f || gs.sloITL < 0.0f || gs.sloTPS < 0.0f)
            blockStatus = WVA_CAND_ERR_CONFIG;
    }
    const size_t candBase = ((size_t)pairLocal * R) * (size_t)B;
    if (blockStatus != WVA_CAND_OK) {
        if (gp.status || gp.cube) {
            const size_t n = (size_t)R * B;
            for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
                if (gp.status) gp.status[candBase + i] = (unsigned char)blockStatus;
                if (gp.cube) { float4 z = make_float4(0, 0, 0, 0); float4* c = reinterpret_cast<float4*>(&gp.cube[candBase + i]); c[0] = z; c[1] = z; }
            }
        }
        return;
    }
    __syncthreads();
    ServFormula sf; sf.init(gs.sp, gs.inTok, gs.outTok);
    double2* gtab = gp.pair_tab + (size_t)pairSlice * B;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        float r = sf.rate(i + 1);
        rateF[i] = r;
        double d = (double)r;
        double y = rcp_refined(d);
        rateD[i] = d; rcp[i] = y;
        gtab[i] = make_double2(d, y);
        if (!(r > 0.0f) || !(r < CUDART_INF_F)) atomicMin(&sh_nGood, i);
    }
    const bool tame = tame_parms(gs.sp, gs.inTok, gs.outTok);
    __syncthreads();
    const int nGood = sh_nGood;

    unsigned long long bestKey = WVA_KEY_NONE;
    float bestItl = 0.0f, bestTtft = 0.0f, bestRho = 0.0f;
    unsigned long long steps = 0, algSteps = 0, okCount = 0;

    for (int r = threadIdx.x + 1; r <= R; r += blockDim.x) {
        const float rate = gs.totalRate / (float)r;
        const float lambda = rate / 1000.0f;
        double lam = (double)lambda;
        const bool lamOk = (lam >= 0x1p-100 && lam <= 0x1p20);
        lam = pin(lam);
        const float cost = gs.accCost * (float)go_muli(gs.numInst, (long long)r);
        float value = transition_penalty(gs.curAcc, gs.curRep, gs.curCost, a, (long long)r, cost);
        value = value + 0.0f;
        // shared ramp state
        double p = 1.0, sum = 1.0, uN = 0.0, dn = 0.0;
        double exInSys = 0.0, exSumP = 0.0;     // exact normalised sums of a stopped row
        unsigned thrHi = 0u, hmin = 0x3ff00000u;
        bool stopped = false;        // ramp died out: sums frozen
        bool broken = !lamOk;        // chain left the value window (or bad table entry): rest of the row goes to the exact kernels
        const size_t rowBase = candBase + (size_t)(r - 1) * B;
        for (int n = 0; n < B; ++n) {
            const int b = n + 1;
            // ---- one step of the shared ramp: state b ----
            if (!stopped && !broken) {
                if (b > nGood) broken = true;
                else {
                    const double t = p * lam;
                    const double pn = div_core(t, rateD[n], rcp[n]);
                    const unsigned hq = (unsigned)__double2hiint(pn);
                    if (hq - WVA_WIN_LO >= WVA_WIN_SPAN) broken = true;
                    else {
                        sum += pn; dn += 1.0; uN += dn * pn; p = pn;
                        ++steps;
                        if (n == 0 && pn >= 0x1p-400)      // 2^-68 min(1,p1) / K_max: the neglected mass is < 2^-58 of every aggregate
                            thrHi = (unsigned)__double2hiint((0x1p-68 * fmin(1.0, pn)) / (double)(11 * B));
                        hmin = hq < hmin ? hq : hmin;
                        if (hq < thrHi && tame && lambda <= 0.998f * rateF[n] && sum <= 0x1p400) {
                            // The chain has died out (solve_stream's truncation rule with a 2^10 stricter threshold):
                            // every candidate b >= this state sees the SAME normalised prefix, exactly.  Run the
                            // reference's second pass once for the row (mm1modelstatedependent.go:49-55 up to here).
                            stopped = true;
                            const double S = sum;
                            if ((int)(hmin >> 20) - (int)((unsigned)__double2hiint(S) >> 20) < -1000) broken = true;
                            else {
                                const double yS = rcp_refined(S);
                                double q = div_core(1.0, S, yS), pp = 1.0, di = 0.0;
                                exSumP = q; exInSys = 0.0;
                                for (int i = 1; i <= b; ++i) {
                                    pp = div_core(pp * lam, rateD[i - 1], rcp[i - 1]);
                                    q = div_core(pp, S, yS);
                                    di += 1.0;
                                    exInSys += di * q;
                                    exSumP += q;
                                }
                                steps += (unsigned long long)b;
                            }
                        }
                    }
                }
            }
            // ---- candidate (r, b) ----
            const size_t ci = rowBase + (size_t)n;
            const int K = 11 * b;
            const float lambdaMax = rateF[n] * (1.0f - WVA_EPSILON);
            const float rateMax = lambdaMax * 1000.0f;
            int st;
            bool feasible = false;
            wva_metrics m;
            m.throughput = m.avg_resp_time = m.avg_wait_time = m.avg_num_in_serv = 0.0f;
            m.avg_prefill_time = m.avg_token_time = m.max_rate = m.rho = 0.0f;
            if (b > nGood) {                                   // bad table entry: literal path decides
                int k = atomicAdd(gp.slow_count, 1);
                if (k < gp.slow_cap) gp.slow_list[k] = (unsigned long long)ci;
                continue;
            }
            if (rate <= 0.0f) st = WVA_CAND_ERR_RATE_LE0;
            else if (rate > rateMax) st = WVA_CAND_ERR_RATE_MAX;
            else if (lambda < 0.0f) st = WVA_CAND_ERR_MODEL;
            else {
                SolveStats so;
                bool certified = false;
                if (!broken) {
                    if (stopped) {
                        // exact: avgNumInServers is captured at i == b (mm1modelstatedependent.go:52-54) from sums that no
                        // longer change; float32(p[K]) < 2^-58 so throughput == lambda
                        const double inServ = exInSys + (1.0 - exSumP) * (double)b;
                        finish_stats(so, lambda, inServ, exInSys, 0.0f);
                        certified = true;
                    } else {
                        CertIn c; c.pN = p; c.sumRamp = sum; c.uN = uN; c.lam = lam; c.sTail = rateD[n]; c.N = b; c.K = K; c.lambda = lambda;
                        certified = cert_eval(c, so);
                    }
                }
                if (!certified) {
                    // exact chain in k_grid_list; when that list is full, right here
                    int k = atomicAdd(gp.heavy_count, 1);
                    if (k < gp.heavy_cap) { gp.heavy_list[k] = (unsigned long long)ci; gp.heavy_cost[k] = (float)K; continue; }
                    ServTable tb; tb.rateF = rateF; tb.rateD = rateD; tb.rcp = rcp;
                    float rt, dc;
                    int st2 = analyze_table(tb, gs, b, rate, tame, 0, m, rt, steps, dc);
                    if (st2 < 0) { int k2 = atomicAdd(gp.slow_count, 1); if (k2 < gp.slow_cap) gp.slow_list[k2] = (unsigned long long)ci; continue; }
                    so.avgServTime = 0.0f;      // metrics already final in m
                    st = st2;
                    goto have_metrics;
                }
                st = WVA_CAND_OK;
                {
                    const float effConc = effective_concurrency(so.avgServTime, gs.sp, gs.inTok, gs.outTok, b);
                    float rho = so.avgNumInServers / (float)b;
                    rho = go_minf(go_maxf(rho, 0.0f), 1.0f);
                    m.throughput = so.throughput * 1000.0f;
                    m.avg_resp_time = so.avgRespTime;
                    m.avg_wait_time = so.avgWaitTime;
                    m.avg_num_in_serv = so.avgNumInServers;
                    m.avg_prefill_time = prefill_time(gs.sp, gs.inTok, effConc);
                    m.avg_token_time = decode_time(gs.sp, effConc);
                    m.max_rate = rateMax;
                    m.rho = rho;
                }
            have_metrics:
                if (st != WVA_CAND_OK) {
                    m.throughput = m.avg_resp_time = m.avg_wait_time = m.avg_num_in_serv = 0.0f;
                    m.avg_prefill_time = m.avg_token_time = m.max_rate = m.rho = 0.0f;
                    goto write_out;
                }
                okCount++;
                algSteps += 2ULL * (unsigned long long)(K + 1);
                const float lamMaxBack = rateMax / 1000.0f;
                const float rateTPS = (lamMaxBack * (1.0f - WVA_STABILITY_SAFETY)) * 1000.0f;
                const float ttft = m.avg_wait_time + m.avg_prefill_time;
                const float itl = m.avg_token_time;
                feasible = (!(gs.sloTTFT > 0.0f) || ttft <= gs.sloTTFT) && (!(gs.sloITL > 0.0f) || itl <= gs.sloITL) &&
                           (!(gs.sloTPS > 0.0f) || rate <= rateTPS) && (r >= gs.minReplicas);
                if (feasible && value == value) {              // a NaN value is never selected
                    const unsigned long long key = make_key(value, a, r, b);
                    if (key < bestKey) { bestKey = key; bestItl = itl; bestTtft = ttft; bestRho = m.rho; }
                }
            }
        write_out:
            if (gp.cube) {
                float4* c = reinterpret_cast<float4*>(&gp.cube[ci]);
                c[0] = make_float4(m.throughput, m.avg_resp_time, m.avg_wait_time, m.avg_num_in_serv);
                c[1] = make_float4(m.avg_prefill_time, m.avg_token_time, m.max_rate, m.rho);
            }
            if (gp.status) gp.status[ci] = (unsigned char)(st | (feasible ? WVA_CAND_FEASIBLE : 0));
        }
    }
    // ---- block argmin + counters ------------------------------------------------------------------
    unsigned long long warpKey = bestKey;
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, warpKey, o);
        if (other < warpKey) warpKey = other;
        steps += __shfl_down_sync(0xffffffffu, steps, o);
        algSteps += __shfl_down_sync(0xffffffffu, algSteps, o);
        okCount += __shfl_down_sync(0xffffffffu, okCount, o);
    }
    if (lane == 0) {
        if (warpKey != WVA_KEY_NONE) atomicMin(&sh_key, warpKey);
        atomicAdd(&sh_cnt[0], steps); atomicAdd(&sh_cnt[1], algSteps); atomicAdd(&sh_cnt[2], okCount);
    }
    __syncthreads();
    const unsigned long long blockKey = sh_key;
    if (blockKey != WVA_KEY_NONE && bestKey == blockKey) {
        GridSlot sl_; sl_.key = blockKey; sl_.itl = bestItl; sl_.ttft = bestTtft; sl_.rho = bestRho; sl_.sl = sl; sl_.pad = 0;
        const int r = (int)((blockKey >> 14) & 0x3ff) + 1;
        sl_.cost = gs.accCost * (float)go_muli(gs.numInst, (long long)r);
        gp.block_slot[blockIdx.x] = sl_;
        atomicMin(&gp.keys[sl], blockKey);
    }
    if (threadIdx.x == 0) {
        atomicAdd(&gp.counters[0], sh_cnt[0]); atomicAdd(&gp.counters[1], sh_cnt[1]); atomicAdd(&gp.counters[2], sh_cnt[2]);
    }
}

// Candidate evaluated through the formula-based analyzer.  Same arithmetic as analyze_table, hence
// the same bits.  tab: the pair's published {rate, reciprocal} table (unified-loop streaming solver)
// or nullptr; scratch: p[] for the literal path or nullptr.
__device__ int analyze_candidate(const DevSystem& sys, int s, int a, int r, int b, const double2* tab, double* scratch,
                                 GridServer& gs, wva_metrics& m, float& rate, float& rateTPS, int& fault,
                                 unsigned long long& steps) {
    load_grid_server(sys, s, a, gs);
    Analyzer qa;
    qa.build(gs.sp, b, (long long)b * WVA_MAX_QUEUE_TO_BATCH_RATIO, gs.inTok, gs.outTok, scratch);
    qa.tab = tab; qa.uni = true; qa.cert = sys.cert != 0;
    const float lamMaxBack = qa.rateMax / 1000.0f;
    rateTPS = (lamMaxBack * (1.0f - WVA_STABILITY_SAFETY)) * 1000.0f;
    rate = gs.totalRate / (float)r;
    int st = qa.analyze(rate, m);
    fault = qa.fault;
    steps += qa.steps;
    return st;
}

// Candidates from a list, one thread each.
//   scratch == nullptr : the deferred long chains (streaming, unified-loop solver on the published
//                        tables); `order` sorts them by estimated length so that the lanes of a warp
//                        finish together; a chain that needs the materialised path is appended to
//                        the literal list.
//   scratch != nullptr : the literal list (p[] materialised, `stride` doubles per thread).
// Each thread publishes key + metrics in gp.list_slot[slot_base + t] for k_grid_claim.
__global__ void __launch_bounds__(128)
k_grid_list(DevSystem sys, GridParams gp, const unsigned long long* __restrict__ list, const int* __restrict__ order,
            int nList, double* scratch, long long stride, int slot_base) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long steps = 0, alg = 0, okc = 0;
    if (t < nList) {
        const size_t ci = (size_t)list[order ? order[t] : t];
        const int B = gp.b_max;
        const int b = (int)(ci % B) + 1;
        const int r = (int)((ci / B) % gp.r_max) + 1;
        const int pairLocal = (int)(ci / ((size_t)B * gp.r_max));
        const int sl = pairLocal / sys.A, a = pairLocal % sys.A, s = gp.s0 + sl;
        GridServer gs; wva_metrics m; float rate, rateTPS; int fault = 0;
        GridSlot slot; slot.key = WVA_KEY_NONE; slot.cost = slot.itl = slot.ttft = slot.rho = 0.0f; slot.sl = sl; slot.pad = 0;
        int st = analyze_candidate(sys, s, a, r, b, scratch ? nullptr : gp.pair_tab + (size_t)(pairLocal - gp.pair_base) * B,
                                   scratch ? scratch + (size_t)t * stride : nullptr, gs, m, rate, rateTPS, fault, steps);
        if (fault == 1) {
            int k = atomicAdd(gp.slow_count, 1);
            if (k < gp.slow_cap) gp.slow_list[k] = (unsigned long long)ci;
        } else {
            bool feasible = false;
            if (st == WVA_CAND_OK) {
                unsigned long long key = candidate_key(gs, a, r, b, rate, rateTPS, m, feasible);
                if (key != WVA_KEY_NONE) {
                    slot.key = key; slot.itl = m.avg_token_time; slot.ttft = m.avg_wait_time + m.avg_prefill_time;
                    slot.rho = m.rho; slot.cost = gs.accCost * (float)go_muli(gs.numInst, (long long)r);
                    if (key < gp.keys[sl]) atomicMin(&gp.keys[sl], key);
                }
                alg = 2ULL * (unsigned long long)(11 * b + 1);
                okc = 1;
            } else {
                m.throughput = m.avg_resp_time = m.avg_wait_time = m.avg_num_in_serv = 0.0f;
                m.avg_prefill_time = m.avg_token_time = m.max_rate = m.rho = 0.0f;
            }
            if (gp.cube) {
                float4* c = reinterpret_cast<float4*>(&gp.cube[ci]);
                c[0] = make_float4(m.throughput, m.avg_resp_time, m.avg_wait_time, m.avg_num_in_serv);
                c[1] = make_float4(m.avg_prefill_time, m.avg_token_time, m.max_rate, m.rho);
            }
            if (gp.status) gp.status[ci] = (unsigned char)(st | (feasible ? WVA_CAND_FEASIBLE : 0));
        }
        gp.list_slot[slot_base + t] = slot;
    }
    for (int o = 16; o > 0; o >>= 1) {
        steps += __shfl_down_sync(0xffffffffu, steps, o);
        alg += __shfl_down_sync(0xffffffffu, alg, o);
        okc += __shfl_down_sync(0xffffffffu, okc, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (steps) atomicAdd(&gp.counters[0], steps);
        if (alg) atomicAdd(&gp.counters[1], alg);
        if (okc) atomicAdd(&gp.counters[2], okc);
    }
}

// ---- ordering of the deferred chains by estimated length (256 buckets, longest first) --------
__device__ __forceinline__ int heavy_bucket(float cost) {
    int k = (int)(__float_as_uint(cost < 1.0f ? 1.0f : cost) >> 19) - (127 << 4);    // 16 buckets per octave
    k = k < 0 ? 0 : (k > 255 ? 255 : k);
    return 255 - k;
}
__global__ void k_heavy_hist(const float* __restrict__ cost, int n, int* __restrict__ hist) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) atomicAdd(&hist[heavy_bucket(cost[t])], 1);
}
__global__ void k_heavy_prefix(int* hist /*[256] -> exclusive prefix in place*/) {
    __shared__ int sh[256];
    int t = threadIdx.x;
    sh[t] = hist[t];
    __syncthreads();
    if (t == 0) { int run = 0; for (int i = 0; i < 256; ++i) { int c = sh[i]; sh[i] = run; run += c; } }
    __syncthreads();
    hist[t] = sh[t];
}
__global__ void k_heavy_scatter(const float* __restrict__ cost, int n, int* __restrict__ cursor, int* __restrict__ order) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) order[atomicAdd(&cursor[heavy_bucket(cost[t])], 1)] = t;
}

// Per-server winners: keys[] holds the minimum key of every server; the slot that carries that key
// (exactly one: keys are unique per candidate) writes the wva_grid_best record.  No re-evaluation.
__global__ void k_grid_best_init(int ns, wva_grid_best* __restrict__ best) {
    int sl = blockIdx.x * blockDim.x + threadIdx.x;
    if (sl >= ns) return;
    wva_grid_best out;
    out.acc = -1; out.replicas = 0; out.batch = 0; out.cost = out.value = out.itl = out.ttft = out.rho = 0.0f;
    best[sl] = out;
}
__global__ void k_grid_claim(GridParams gp, const GridSlot* __restrict__ slots, int nSlots, wva_grid_best* __restrict__ best) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nSlots) return;
    const GridSlot sl_ = slots[t];
    if (sl_.key == WVA_KEY_NONE) return;
    if (gp.keys[sl_.sl] != sl_.key) return;
    wva_grid_best out;
    out.acc = (int)((sl_.key >> 24) & 0xff); out.replicas = (int)((sl_.key >> 14) & 0x3ff) + 1; out.batch = (int)(sl_.key & 0x3fff) + 1;
    out.cost = sl_.cost;
    out.value = unsortable_f32((unsigned)(sl_.key >> 32));
    out.itl = sl_.itl; out.ttft = sl_.ttft; out.rho = sl_.rho;
    best[sl_.sl] = out;
}

// ---------------------------------------------------------------------------------------
// pkg/analyzer batched public API
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_queue_analyze(int n, const wva_queue_config* __restrict__ cfg, const float* __restrict__ rate,
                wva_metrics* __restrict__ metrics, unsigned char* __restrict__ status, double* scratch,
                const long long* __restrict__ scratch_off, const int* __restrict__ list, int* fault_list, int* fault_count) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = list ? list[t] : t;
    wva_queue_config c = cfg[i];
    wva_metrics m;
    m.throughput = m.avg_resp_time = m.avg_wait_time = m.avg_num_in_serv = 0.0f;
    m.avg_prefill_time = m.avg_token_time = m.max_rate = m.rho = 0.0f;
    int st;
    if (!config_ok(c.max_batch_size, c.max_queue_size, c.avg_input_tokens, c.avg_output_tokens)) st = WVA_CAND_ERR_CONFIG;
    else {
        ServiceParms sp; sp.alpha = c.alpha; sp.beta = c.beta; sp.gamma = c.gamma; sp.delta = c.delta;
        Analyzer qa;
        qa.build(sp, c.max_batch_size, c.max_queue_size, c.avg_input_tokens, c.avg_output_tokens,
                 scratch ? scratch + scratch_off[t] : nullptr);
        st = qa.analyze(rate[i], m);
        if (qa.fault == 1) { fault_list[atomicAdd(fault_count, 1)] = i; return; }
        if (st != WVA_CAND_OK) {
            m.throughput = m.avg_resp_time = m.avg_wait_time = m.avg_num_in_serv = 0.0f;
            m.avg_prefill_time = m.avg_token_time = m.max_rate = m.rho = 0.0f;
        }
    }
    metrics[i] = m;
    status[i] = (unsigned char)st;
}

__global__ void __launch_bounds__(128)
k_queue_size(int n, const wva_queue_config* __restrict__ cfg, const float* __restrict__ target,
             float* __restrict__ rates, wva_metrics* __restrict__ metrics, float* __restrict__ achieved,
             unsigned char* __restrict__ status, double* scratch, const long long* __restrict__ scratch_off,
             const int* __restrict__ list, int* fault_list, int* fault_count) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = list ? list[t] : t;
    wva_queue_config c = cfg[i];
    wva_metrics m;
    m.throughput = m.avg_resp_time = m.avg_wait_time = m.avg_num_in_serv = 0.0f;
    m.avg_prefill_time = m.avg_token_time = m.max_rate = m.rho = 0.0f;
    float rt[3] = {0.0f, 0.0f, 0.0f}, ach[3] = {0.0f, 0.0f, 0.0f};
    bool ok = false;
    if (config_ok(c.max_batch_size, c.max_queue_size, c.avg_input_tokens, c.avg_output_tokens)) {
        ServiceParms sp; sp.alpha = c.alpha; sp.beta = c.beta; sp.gamma = c.gamma; sp.delta = c.delta;
        Analyzer qa;
        qa.build(sp, c.max_batch_size, c.max_queue_size, c.avg_input_tokens, c.avg_output_tokens,
                 scratch ? scratch + scratch_off[t] : nullptr);
        ok = size_queue(qa, target[3 * i], target[3 * i + 1], target[3 * i + 2], rt, m, ach);
        if (qa.fault == 1) { fault_list[atomicAdd(fault_count, 1)] = i; return; }
        if (!ok) {
            m.throughput = m.avg_resp_time = m.avg_wait_time = m.avg_num_in_serv = 0.0f;
            m.avg_prefill_time = m.avg_token_time = m.max_rate = m.rho = 0.0f;
            rt[0] = rt[1] = rt[2] = ach[0] = ach[1] = ach[2] = 0.0f;
        }
    }
    metrics[i] = m;
    for (int k = 0; k < 3; ++k) { rates[3 * i + k] = rt[k]; achieved[3 * i + k] = ach[k]; }
    status[i] = ok ? 0 : 1;
}

// ---------------------------------------------------------------------------------------
// Optimize
// ---------------------------------------------------------------------------------------

// SolveUnlimited, solver.go:63-79: per server argmin of value with strict '<' from MaxFloat32,
// candidates visited in ascending accelerator index (canonical order for Go's random map order).
__global__ void k_solve_unlimited(DevSystem sys, int s0, int ns, DevAllocs pairs, const unsigned char* __restrict__ feasible,
                                  int* __restrict__ chosen_acc, DevAllocs chosen) {
    int sl = blockIdx.x * blockDim.x + threadIdx.x;
    if (sl >= ns) return;
    int s = s0 + sl;
    float minVal = 3.40282346638528859811704183484516925e+38f;
    int minKey = -1;
    for (int a = 0; a < sys.A; ++a) {
        size_t i = (size_t)s * sys.A + a;
        if (!feasible[i]) continue;
        float v = pairs.value[i];
        if (v < minVal) { minVal = v; minKey = a; }
    }
    chosen_acc[s] = minKey;
    store_alloc(chosen, (size_t)s, minKey >= 0 ? load_alloc(pairs, (size_t)s * sys.A + minKey) : empty_alloc());
}

// System.AllocateByType (system.go:271-300) over servers [s0, s0+ns): one thread per accelerator
// type walks the servers in ascending index, so the float32 cost sum has a fixed order.
// totals = { long long count[T]; float cost[T] }.
__global__ void k_totals(DevSystem sys, int s0, int ns, const int* __restrict__ chosen_acc, DevAllocs chosen,
                         long long* __restrict__ count, float* __restrict__ cost) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= sys.T) return;
    long long c = 0; float k = 0.0f;
    for (int s = s0; s < s0 + ns; ++s) {
        if (chosen_acc[s] < 0) continue;
        int gi = chosen.acc[s];
        int m = sys.srv_model[s];
        if (gi < 0 || m < 0) continue;
        if (sys.acc_type[gi] != t) continue;
        c += go_muli(go_muli(chosen.num_replicas[s], num_instances(sys, m, gi)), (long long)sys.acc_multiplicity[gi]);
        k = k + chosen.cost[s];
    }
    count[t] = c; cost[t] = k;
}

// ---- greedy ------------------------------------------------------------------------------

// cmp.Compare for float32: NaN lowest, -0 == +0
__device__ __forceinline__ int go_cmpf(float x, float y) {
    bool xn = x != x, yn = y != y;
    if (xn) return yn ? 0 : -1;
    if (yn) return 1;
    if (x < y) return -1;
    if (x > y) return 1;
    return 0;
}

struct GreedyBufs {
    int* order;            // [S*A] per server: candidate keys sorted by value (stable on accelerator index)
    int* nCand;            // [S]
    int* curIndex;         // [S]
    float* delta;          // [S]
    int* stamp;            // [S] recency stamp (initial -server index)
    int* heap;             // [S] server ids
    int* groupStart;       // [102] population / start offset of each priority (1..100)
    int* groupItems;       // [S]
    int* unalloc;          // [S]
    int* ticketAcc;        // [S] allocateEqually: chosen candidate key
    int* ticketRep;        // [S] replicas handed out
    unsigned char* ticketState;   // [S] 0 absent, 1 present, 2 active
    long long* available;  // [T]
    int* nanFlag;          // [1]
};

// per server: sort candidate keys by value (greedy.go:57-63), initial delta (:64-71)
__global__ void k_greedy_prepare(DevSystem sys, DevAllocs pairs, const unsigned char* __restrict__ feasible, GreedyBufs g) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= sys.S) return;
    int* ord = g.order + (size_t)s * sys.A;
    int n = 0;
    for (int a = 0; a < sys.A; ++a) {
        size_t i = (size_t)s * sys.A + a;
        if (!feasible[i]) continue;
        float v = pairs.value[i];
        if (v != v) atomicExch(g.nanFlag, 1);
        int j = n++;
        // stable insertion sort by cmp.Compare(value)
        while (j > 0 && go_cmpf(pairs.value[(size_t)s * sys.A + ord[j - 1]], v) > 0) { ord[j] = ord[j - 1]; --j; }
        ord[j] = a;
    }
    g.nCand[s] = n;
    g.curIndex[s] = 0;
    g.stamp[s] = -s;
    float d = 0.0f;
    if (n > 1) d = pairs.value[(size_t)s * sys.A + ord[1]] - pairs.value[(size_t)s * sys.A + ord[0]];
    else if (n == 1) d = 3.40282346638528859811704183484516925e+38f;
    g.delta[s] = d;
}

struct GreedyCtx {
    DevSystem sys; DevAllocs pairs; GreedyBufs g; int* chosen;
    __device__ __forceinline__ float curValue(int s) const {
        return pairs.value[(size_t)s * sys.A + g.order[(size_t)s * sys.A + g.curIndex[s]]];
    }
    // orderFunc, greedy.go:76-85, refined to a strict order by the recency stamp: an entry
    // re-inserted by slices.BinarySearchFunc lands before every equal entry (largest stamp first);
    // the initial stable sort keeps equal entries in ascending server index (stamp = -index).
    __device__ __forceinline__ bool before(int x, int y) const {
        int px = sys.srv_priority[x], py = sys.srv_priority[y];
        if (px != py) return px < py;
        float dx = g.delta[x], dy = g.delta[y];
        int c;
        if (dx == dy) c = go_cmpf(curValue(y), curValue(x));
        else c = go_cmpf(dy, dx);
        if (c != 0) return c < 0;
        return g.stamp[x] > g.stamp[y];
    }
    __device__ void siftDown(int* h, int n, int i) const {
        int v = h[i];
        for (;;) {
            int l = 2 * i + 1;
            if (l >= n) break;
            int r = l + 1;
            int c = (r < n && before(h[r], h[l])) ? r : l;
            if (!before(h[c], v)) break;
            h[i] = h[c]; i = c;
        }
        h[i] = v;
    }
    __device__ void siftUp(int* h, int i) const {
        int v = h[i];
        while (i > 0) {
            int p = (i - 1) >> 1;
            if (!before(v, h[p])) break;
            h[i] = h[p]; i = p;
        }
        h[i] = v;
    }
    __device__ __forceinline__ long long unitsPerReplica(int s, int gi) const {
        return go_muli(num_instances(sys, sys.srv_model[s], gi), (long long)sys.acc_multiplicity[gi]);
    }
    __device__ __forceinline__ int candKey(int s, int idx) const { return g.order[(size_t)s * sys.A + idx]; }
};

// allocate, greedy.go:107-166, over the servers in items[0..n); returns number of unallocated
// entries appended to g.unalloc (starting at unallocBase).
__device__ int greedy_allocate(GreedyCtx& c, const int* items, int n, int unallocBase) {
    int* h = c.g.heap;
    int hn = 0;
    for (int i = 0; i < n; ++i) if (c.g.nCand[items[i]] > 0) h[hn++] = items[i];
    for (int i = hn / 2 - 1; i >= 0; --i) c.siftDown(h, hn, i);
    int stampCounter = 1;
    int nUn = 0;
    while (hn > 0) {
        int s = h[0];
        h[0] = h[--hn];
        if (hn > 0) c.siftDown(h, hn, 0);
        if (c.sys.srv_model[s] < 0) continue;
        int key = c.candKey(s, c.g.curIndex[s]);
        size_t ai = (size_t)s * c.sys.A + key;
        int gi = c.pairs.acc[ai];
        if (gi < 0) continue;                                       // GetAccelerator("") == nil
        int t = c.sys.acc_type[gi];
        long long count = go_muli(c.pairs.num_replicas[ai], c.unitsPerReplica(s, gi));
        if (c.g.available[t] >= count) {
            c.g.available[t] -= count;
            c.chosen[s] = key;
        } else {
            int ci = ++c.g.curIndex[s];
            int len = c.g.nCand[s];
            if (ci + 1 < len) {
                c.g.delta[s] = c.pairs.value[(size_t)s * c.sys.A + c.candKey(s, ci + 1)] -
                               c.pairs.value[(size_t)s * c.sys.A + c.candKey(s, ci)];
            } else if (ci == len) {
                c.g.unalloc[unallocBase + nUn++] = s;
                continue;
            } else {
                c.g.delta[s] = 3.40282346638528859811704183484516925e+38f;
            }
            c.g.stamp[s] = stampCounter++;
            h[hn] = s;
            c.siftUp(h, hn);
            ++hn;
        }
    }
    return nUn;
}

// allocateMaximally, greedy.go:194-223
__device__ void greedy_allocate_maximally(GreedyCtx& c, const int* list, int n) {
    for (int i = 0; i < n; ++i) {
        int s = list[i];
        if (c.sys.srv_model[s] < 0) continue;
        for (int k = 0; k < c.g.nCand[s]; ++k) {
            int key = c.candKey(s, k);
            size_t ai = (size_t)s * c.sys.A + key;
            int gi = c.pairs.acc[ai];
            if (gi < 0) continue;
            long long upr = c.unitsPerReplica(s, gi);
            if (upr <= 0) continue;
            int t = c.sys.acc_type[gi];
            long long cur = c.pairs.num_replicas[ai];
            long long maxRep = go_divi(c.g.available[t], upr);
            if (cur < maxRep) maxRep = cur;
            if (maxRep > 0) {
                float factor = (float)maxRep / (float)cur;
                c.pairs.cost[ai] = c.pairs.cost[ai] * factor;
                c.pairs.value[ai] = c.pairs.value[ai] * factor;
                c.pairs.num_replicas[ai] = maxRep;
                c.chosen[s] = key;
                c.g.available[t] -= go_muli(maxRep, upr);
                break;
            }
        }
    }
}

// allocateEqually, greedy.go:239-316
__device__ void greedy_allocate_equally(GreedyCtx& c, const int* list, int n) {
    int live = 0;
    for (int i = 0; i < n; ++i) {
        int s = list[i];
        c.g.ticketRep[s] = 0; c.g.ticketAcc[s] = -1;
        if (c.sys.srv_model[s] < 0) { c.g.ticketState[s] = 0; continue; }
        c.g.ticketState[s] = 1; ++live;
    }
    while (live > 0) {
        for (int i = 0; i < n; ++i) {
            int s = list[i];
            unsigned char stt = c.g.ticketState[s];
            if (stt == 0) continue;
            if (stt == 1) {
                bool found = false;
                for (int k = 0; k < c.g.nCand[s]; ++k) {
                    int key = c.candKey(s, k);
                    int gi = c.pairs.acc[(size_t)s * c.sys.A + key];
                    if (gi < 0) continue;
                    long long upr = c.unitsPerReplica(s, gi);
                    if (upr > 0 && c.g.available[c.sys.acc_type[gi]] >= upr) { c.g.ticketAcc[s] = key; found = true; break; }
                }
                if (!found) { c.g.ticketState[s] = 0; --live; continue; }
                c.g.ticketState[s] = 2;
            }
            int key = c.g.ticketAcc[s];
            size_t ai = (size_t)s * c.sys.A + key;
            int gi = c.pairs.acc[ai];
            int t = c.sys.acc_type[gi];
            long long upr = c.unitsPerReplica(s, gi);
            long long avail = go_divi(c.g.available[t], upr);
            long long cur = c.pairs.num_replicas[ai];
            long long allocatable = avail < cur ? avail : cur;
            if (allocatable > 0) {
                c.g.ticketRep[s]++;
                c.g.available[t] -= upr;
            } else {
                c.g.ticketState[s] = 0; --live;
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        int s = list[i];
        int got = c.g.ticketRep[s];
        if (got <= 0) continue;
        int key = c.g.ticketAcc[s];
        size_t ai = (size_t)s * c.sys.A + key;
        long long cur = c.pairs.num_replicas[ai];
        float factor = (float)got / (float)cur;
        c.pairs.cost[ai] = c.pairs.cost[ai] * factor;
        c.pairs.value[ai] = c.pairs.value[ai] * factor;
        c.pairs.num_replicas[ai] = got;
        c.chosen[s] = key;
    }
}

// bestEffort, greedy.go:169-190 (list is grouped by priority already)
__device__ void greedy_best_effort(GreedyCtx& c, const int* list, int n, int policy) {
    if (policy == WVA_POLICY_PRIORITY_EXHAUSTIVE) greedy_allocate_maximally(c, list, n);
    else if (policy == WVA_POLICY_ROUND_ROBIN) greedy_allocate_equally(c, list, n);
    else if (policy == WVA_POLICY_PRIORITY_ROUND_ROBIN) {
        int i = 0;
        while (i < n) {                                              // makePriorityGroups, :321-341
            int j = i + 1;
            int pr = c.sys.srv_priority[list[i]];
            while (j < n && c.sys.srv_priority[list[j]] == pr) ++j;
            greedy_allocate_equally(c, list + i, j - i);
            i = j;
        }
    }
}

// Bucket servers by priority (counting sort).  Priorities come from Server.Priority() and are in
// [1,100] (serviceclass.go:28-37, server.go:92-97); wva_system_upload rejects anything else.
__global__ void k_greedy_bucket_count(DevSystem sys, GreedyBufs g) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= sys.S) return;
    atomicAdd(&g.groupStart[sys.srv_priority[s]], 1);
}

// SolveGreedy, greedy.go:35-104: the sequential assignment.  One thread: every step depends on
// the capacities left by the previous one.  Servers are bucketed by priority; inside a bucket the
// binary heap reproduces the sorted-slice order of the reference (see GreedyCtx::before).
__global__ void k_greedy_solve(DevSystem sys, DevAllocs pairs, GreedyBufs g, int* chosen, int delayedBestEffort, int policy) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    GreedyCtx c; c.sys = sys; c.pairs = pairs; c.g = g; c.chosen = chosen;
    for (int t = 0; t < sys.T; ++t) g.available[t] = sys.type_capacity[t];
    // groupStart[p] holds the population of priority p (1..100): exclusive prefix sum, then a
    // stable scatter in ascending server index
    int cursor[102];
    int run = 0;
    for (int p = 0; p <= 101; ++p) { int cnt = g.groupStart[p]; g.groupStart[p] = run; cursor[p] = run; run += cnt; }
    for (int s = 0; s < sys.S; ++s) g.groupItems[cursor[sys.srv_priority[s]]++] = s;
    if (delayedBestEffort) {
        // one allocate() over everything (the comparator orders by priority first), then one bestEffort
        int nUn = greedy_allocate(c, g.groupItems, sys.S, 0);
        greedy_best_effort(c, g.unalloc, nUn, policy);
    } else {
        for (int p = 1; p <= 100; ++p) {                              // makePriorityGroups(entries), :96-103
            int lo = g.groupStart[p], hi = g.groupStart[p + 1];
            if (hi <= lo) continue;
            int nUn = greedy_allocate(c, g.groupItems + lo, hi - lo, 0);
            greedy_best_effort(c, g.unalloc, nUn, policy);
        }
    }
}

// copy the chosen candidates out (after best-effort scaling)
__global__ void k_greedy_collect(DevSystem sys, DevAllocs pairs, const int* __restrict__ chosen_key, int* __restrict__ chosen_acc,
                                 DevAllocs chosen) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= sys.S) return;
    int key = chosen_key[s];
    chosen_acc[s] = key;
    store_alloc(chosen, (size_t)s, key >= 0 ? load_alloc(pairs, (size_t)s * sys.A + key) : empty_alloc());
}

}  // namespace wva
