// Package native is the cgo binding of the B200 library (include/wva_b200.h).
//
// WRITTEN WITHOUT A GO TOOLCHAIN: the build image of this repository has no `go`, so this file has
// never been compiled.  It is the binding a maintainer drops into the reference tree (module
// github.com/llm-d-incubation/workload-variant-autoscaler) next to the replacement packages in
// go/pkg; build with CGO_ENABLED=1 (the reference Dockerfile:25 sets 0) and
//   CGO_CFLAGS=-I<repo>/include  CGO_LDFLAGS="-L<repo>/inferno-autoscaler_b200 -lwva_b200"
//
// cgo pointer rules: every call passes Go-allocated slices for the duration of the call only; the
// library copies what it needs (wva_b200.h "Conventions") and retains no Go pointer.
package native

/*
#cgo LDFLAGS: -lwva_b200
#include <stdlib.h>
#include "wva_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"
)

// SystemImage is the structure-of-arrays form of config.SystemSpec (wva_system_soa).
type SystemImage struct {
	S, A, M, T int

	AccCost         []float32
	AccMultiplicity []int32
	AccType         []int32
	TypeCapacity    []int64

	PerfAlpha, PerfBeta, PerfGamma, PerfDelta []float32
	PerfMaxBatch, PerfAtTokens, PerfAccCount  []int32
	PerfValid                                 []uint8

	SrvModel                         []int32
	SrvArrivalRPM                    []float32
	SrvInTokens, SrvOutTokens        []int32
	SrvSloTTFT, SrvSloITL, SrvSloTPS []float32
	SrvTargetValid                   []uint8
	SrvPriority, SrvMinReplicas      []int32
	SrvMaxBatch                      []int32
	SrvKeepAcc                       []uint8
	SrvCurAcc, SrvCurReplicas        []int32
	SrvCurCost                       []float32
}

// Allocs mirrors wva_alloc_soa (one record per index).
type Allocs struct {
	Acc                                   []int32
	NumReplicas, BatchSize                []int64
	Cost, Value, ITL, TTFT, Rho, MaxArrv []float32
}

func newAllocs(n int) *Allocs {
	return &Allocs{Acc: make([]int32, n), NumReplicas: make([]int64, n), BatchSize: make([]int64, n),
		Cost: make([]float32, n), Value: make([]float32, n), ITL: make([]float32, n), TTFT: make([]float32, n),
		Rho: make([]float32, n), MaxArrv: make([]float32, n)}
}

func (a *Allocs) c() C.wva_alloc_soa {
	return C.wva_alloc_soa{
		acc:          (*C.int32_t)(unsafe.Pointer(&a.Acc[0])),
		num_replicas: (*C.int64_t)(unsafe.Pointer(&a.NumReplicas[0])),
		batch_size:   (*C.int64_t)(unsafe.Pointer(&a.BatchSize[0])),
		cost:         (*C.float)(unsafe.Pointer(&a.Cost[0])), value: (*C.float)(unsafe.Pointer(&a.Value[0])),
		itl: (*C.float)(unsafe.Pointer(&a.ITL[0])), ttft: (*C.float)(unsafe.Pointer(&a.TTFT[0])),
		rho:                        (*C.float)(unsafe.Pointer(&a.Rho[0])),
		max_arrv_rate_per_replica: (*C.float)(unsafe.Pointer(&a.MaxArrv[0])),
	}
}

// Context wraps wva_ctx: one per process (creation builds the CUDA context), one call at a time —
// the same constraint the reference has through its package globals (pkg/core/system.go:12).
type Context struct {
	mu  sync.Mutex
	ctx *C.wva_ctx
}

var (
	defaultOnce sync.Once
	defaultCtx  *Context
	defaultErr  error
)

// Default returns the process-wide context on CUDA device 0 (WVA_B200_DEVICE overrides in the shim).
func Default() (*Context, error) {
	defaultOnce.Do(func() { defaultCtx, defaultErr = NewContext(0) })
	return defaultCtx, defaultErr
}

func NewContext(device int) (*Context, error) {
	var c *C.wva_ctx
	if rc := C.wva_ctx_create(C.int(device), &c); rc != C.WVA_OK {
		return nil, fmt.Errorf("wva_ctx_create: %d: %s", int(rc), C.GoString(C.wva_last_error(nil)))
	}
	ctx := &Context{ctx: c}
	runtime.SetFinalizer(ctx, func(x *Context) { C.wva_ctx_destroy(x.ctx) })
	return ctx, nil
}

func (c *Context) err(rc C.int, what string) error {
	if rc == C.WVA_OK {
		return nil
	}
	return fmt.Errorf("%s: %d: %s", what, int(rc), C.GoString(C.wva_last_error(c.ctx)))
}

func f32p(s []float32) *C.float { if len(s) == 0 { return nil }; return (*C.float)(unsafe.Pointer(&s[0])) }
func i32p(s []int32) *C.int32_t { if len(s) == 0 { return nil }; return (*C.int32_t)(unsafe.Pointer(&s[0])) }
func i64p(s []int64) *C.int64_t { if len(s) == 0 { return nil }; return (*C.int64_t)(unsafe.Pointer(&s[0])) }
func u8p(s []uint8) *C.uint8_t  { if len(s) == 0 { return nil }; return (*C.uint8_t)(unsafe.Pointer(&s[0])) }

// Upload replaces core.NewSystem + System.SetFromSpec on the native side (wva_system_upload).
func (c *Context) Upload(img *SystemImage) error {
	c.mu.Lock()
	defer c.mu.Unlock()
	h := C.wva_system_soa{
		n_servers: C.int32_t(img.S), n_accels: C.int32_t(img.A), n_models: C.int32_t(img.M), n_types: C.int32_t(img.T),
		acc_cost: f32p(img.AccCost), acc_multiplicity: i32p(img.AccMultiplicity), acc_type: i32p(img.AccType),
		type_capacity: i64p(img.TypeCapacity),
		perf_alpha:    f32p(img.PerfAlpha), perf_beta: f32p(img.PerfBeta), perf_gamma: f32p(img.PerfGamma), perf_delta: f32p(img.PerfDelta),
		perf_max_batch: i32p(img.PerfMaxBatch), perf_at_tokens: i32p(img.PerfAtTokens), perf_acc_count: i32p(img.PerfAccCount),
		perf_valid: u8p(img.PerfValid),
		srv_model:  i32p(img.SrvModel), srv_arrival_rpm: f32p(img.SrvArrivalRPM), srv_in_tokens: i32p(img.SrvInTokens),
		srv_out_tokens: i32p(img.SrvOutTokens), srv_slo_ttft: f32p(img.SrvSloTTFT), srv_slo_itl: f32p(img.SrvSloITL),
		srv_slo_tps: f32p(img.SrvSloTPS), srv_target_valid: u8p(img.SrvTargetValid), srv_priority: i32p(img.SrvPriority),
		srv_min_replicas: i32p(img.SrvMinReplicas), srv_max_batch: i32p(img.SrvMaxBatch), srv_keep_acc: u8p(img.SrvKeepAcc),
		srv_cur_acc: i32p(img.SrvCurAcc), srv_cur_replicas: i32p(img.SrvCurReplicas), srv_cur_cost: f32p(img.SrvCurCost),
	}
	return c.err(C.wva_system_upload(c.ctx, &h), "wva_system_upload")
}

// AnalyzePairs replaces Server.Calculate for all servers: S*A records + feasible flags.
func (c *Context) AnalyzePairs(s, a int) (*Allocs, []uint8, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	n := s * a
	if n == 0 {
		return newAllocs(0), nil, nil
	}
	out, fe := newAllocs(n), make([]uint8, n)
	co := out.c()
	if err := c.err(C.wva_analyze_pairs(c.ctx, &co, u8p(fe)), "wva_analyze_pairs"); err != nil {
		return nil, nil, err
	}
	return out, fe, nil
}

// Solve replaces solver.Solver.Solve: chosen candidate key per server (-1 = none) and a copy of it.
func (c *Context) Solve(s int, unlimited, delayedBestEffort bool, policy int) ([]int32, *Allocs, int64, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	b := func(v bool) C.int32_t { if v { return 1 }; return 0 }
	spec := C.wva_optimizer_spec{unlimited: b(unlimited), delayed_best_effort: b(delayedBestEffort), saturation_policy: C.int32_t(policy)}
	if s == 0 {
		return nil, newAllocs(0), 0, nil
	}
	key, out := make([]int32, s), newAllocs(s)
	co := out.c()
	if err := c.err(C.wva_solve(c.ctx, &spec, i32p(key), &co), "wva_solve"); err != nil {
		return nil, nil, 0, err
	}
	return key, out, int64(C.wva_solution_time_usec(c.ctx)), nil
}

// AllocateByType replaces System.AllocateByType (per accelerator type totals of this process' shard).
func (c *Context) AllocateByType(t int) ([]int64, []float32, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	count, cost := make([]int64, t), make([]float32, t)
	if t == 0 {
		return count, cost, nil
	}
	return count, cost, c.err(C.wva_allocate_by_type(c.ctx, i64p(count), f32p(cost)), "wva_allocate_by_type")
}

// Analyze runs Server.Calculate for every pair of the shard and the candidate sweep side by side
// (wva_analyze); results stay on the device for Solve / PairsFetch / GridFetch.
func (c *Context) Analyze(rMax, bMax int) error {
	c.mu.Lock()
	defer c.mu.Unlock()
	return c.err(C.wva_analyze(c.ctx, C.int32_t(rMax), C.int32_t(bMax), 0), "wva_analyze")
}

// SetShard restricts the following calls to servers [first, first+count) (one process per GPU).
func (c *Context) SetShard(first, count int) error {
	c.mu.Lock()
	defer c.mu.Unlock()
	return c.err(C.wva_set_shard(c.ctx, C.int32_t(first), C.int32_t(count)), "wva_set_shard")
}

// TypeTotalsMerge sums the per-rank totals blocks that the host all-gathered (device memory,
// rank-major, 12*T bytes each) in rank order into this context's totals buffer: the exchange step of
// a sharded reconcile.  The collective itself belongs to the caller (NCCL / MPI binding of its choice).
func (c *Context) TypeTotalsMerge(gatheredDev unsafe.Pointer, nRanks int) error {
	c.mu.Lock()
	defer c.mu.Unlock()
	return c.err(C.wva_type_totals_merge(c.ctx, gatheredDev, C.int32_t(nRanks)), "wva_type_totals_merge")
}

// GridBest is wva_grid_best.
type GridBest struct {
	Acc, Replicas, Batch       int32
	Cost, Value, ITL, TTFT, Rho float32
}

// AnalyzeGrid runs the (server x accelerator x replicas x batch) candidate sweep; one winner per server.
func (c *Context) AnalyzeGrid(s, rMax, bMax int) ([]GridBest, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	best := make([]GridBest, s)
	if s == 0 {
		return best, nil
	}
	rc := C.wva_analyze_grid(c.ctx, C.int32_t(rMax), C.int32_t(bMax), (*C.wva_grid_best)(unsafe.Pointer(&best[0])), nil, nil)
	return best, c.err(rc, "wva_analyze_grid")
}

// QueueConfig is wva_queue_config; Metrics is wva_metrics (analyzer.AnalysisMetrics).
type QueueConfig struct {
	MaxBatchSize, MaxQueueSize int32
	Alpha, Beta, Gamma, Delta  float32
	AvgInputTokens, AvgOutputTokens int32
}
type Metrics struct {
	Throughput, AvgRespTime, AvgWaitTime, AvgNumInServ, AvgPrefillTime, AvgTokenTime, MaxRate, Rho float32
}

// QueueAnalyze: n independent QueueAnalyzer.Analyze calls.
func (c *Context) QueueAnalyze(cfg []QueueConfig, rate []float32) ([]Metrics, []uint8, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	n := len(cfg)
	m, st := make([]Metrics, n), make([]uint8, n)
	if n == 0 {
		return m, st, nil
	}
	rc := C.wva_queue_analyze(c.ctx, C.int32_t(n), (*C.wva_queue_config)(unsafe.Pointer(&cfg[0])), f32p(rate),
		(*C.wva_metrics)(unsafe.Pointer(&m[0])), u8p(st))
	return m, st, c.err(rc, "wva_queue_analyze")
}

// QueueSize: n independent QueueAnalyzer.Size calls; target/rates/achieved are (TTFT, ITL, TPS) triples.
func (c *Context) QueueSize(cfg []QueueConfig, target []float32) (rates []float32, m []Metrics, achieved []float32, status []uint8, err error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	n := len(cfg)
	rates, achieved, m, status = make([]float32, 3*n), make([]float32, 3*n), make([]Metrics, n), make([]uint8, n)
	if n == 0 {
		return
	}
	rc := C.wva_queue_size(c.ctx, C.int32_t(n), (*C.wva_queue_config)(unsafe.Pointer(&cfg[0])), f32p(target), f32p(rates),
		(*C.wva_metrics)(unsafe.Pointer(&m[0])), f32p(achieved), u8p(status))
	err = c.err(rc, "wva_queue_size")
	return
}
