// Package analyzer — the public surface of the reference's pkg/analyzer (queueanalyzer.go) over the
// B200 library: NewQueueAnalyzer / Analyze / Size with the reference's types and error behaviour.
// Each call is one native batch of size 1; batch many analyzers through native.Context directly.
// WRITTEN WITHOUT A GO TOOLCHAIN; never compiled (INTEGRATION.md).
package analyzer

import (
	"fmt"

	"github.com/llm-d-incubation/workload-variant-autoscaler/go/internal/native"
)

const Epsilon = float32(0.001)
const StabilitySafetyFraction = float32(0.1)

type PrefillParms struct{ Gamma, Delta float32 }
type DecodeParms struct{ Alpha, Beta float32 }
type ServiceParms struct {
	Prefill *PrefillParms
	Decode  *DecodeParms
}
type Configuration struct {
	MaxBatchSize, MaxQueueSize int
	ServiceParms               *ServiceParms
}
type RequestSize struct{ AvgInputTokens, AvgOutputTokens int }
type RateRange struct{ Min, Max float32 }
type AnalysisMetrics struct {
	Throughput, AvgRespTime, AvgWaitTime, AvgNumInServ, AvgPrefillTime, AvgTokenTime, MaxRate, Rho float32
}
type TargetPerf struct{ TargetTTFT, TargetITL, TargetTPS float32 }
type TargetRate struct{ RateTargetTTFT, RateTargetITL, RateTargetTPS float32 }

type QueueAnalyzer struct {
	MaxBatchSize, MaxQueueSize int
	ServiceParms               *ServiceParms
	RequestSize                *RequestSize
	cfg                        native.QueueConfig
}

func NewQueueAnalyzer(c *Configuration, rq *RequestSize) (*QueueAnalyzer, error) {
	if c.MaxBatchSize <= 0 || c.MaxQueueSize < 0 || c.ServiceParms == nil || c.ServiceParms.Prefill == nil || c.ServiceParms.Decode == nil {
		return nil, fmt.Errorf("invalid configuration {maxBatch=%d, maxQueue=%d}", c.MaxBatchSize, c.MaxQueueSize)
	}
	if rq.AvgInputTokens < 0 || rq.AvgOutputTokens < 1 {
		return nil, fmt.Errorf("invalid request size {inTokens=%d, outTokens=%d}", rq.AvgInputTokens, rq.AvgOutputTokens)
	}
	return &QueueAnalyzer{MaxBatchSize: c.MaxBatchSize, MaxQueueSize: c.MaxQueueSize, ServiceParms: c.ServiceParms, RequestSize: rq,
		cfg: native.QueueConfig{MaxBatchSize: int32(c.MaxBatchSize), MaxQueueSize: int32(c.MaxQueueSize),
			Alpha: c.ServiceParms.Decode.Alpha, Beta: c.ServiceParms.Decode.Beta, Gamma: c.ServiceParms.Prefill.Gamma,
			Delta: c.ServiceParms.Prefill.Delta, AvgInputTokens: int32(rq.AvgInputTokens), AvgOutputTokens: int32(rq.AvgOutputTokens)}}, nil
}

func metricsOf(m native.Metrics) *AnalysisMetrics {
	return &AnalysisMetrics{m.Throughput, m.AvgRespTime, m.AvgWaitTime, m.AvgNumInServ, m.AvgPrefillTime, m.AvgTokenTime, m.MaxRate, m.Rho}
}

// Analyze (queueanalyzer.go:134-174).  NOTE: a fresh model state per call; the stale-rho validity quirk of
// queuemodel.go:30 can only be observed with MaxBatchSize+MaxQueueSize < 2, which the native side reports
// as an invalid model on every call.
func (qa *QueueAnalyzer) Analyze(requestRate float32) (*AnalysisMetrics, error) {
	ctx, err := native.Default()
	if err != nil {
		return nil, err
	}
	m, st, err := ctx.QueueAnalyze([]native.QueueConfig{qa.cfg}, []float32{requestRate})
	if err != nil {
		return nil, err
	}
	if st[0] != 0 {
		return nil, fmt.Errorf("invalid request rate %v (status %d)", requestRate, st[0])
	}
	return metricsOf(m[0]), nil
}

// Size (queueanalyzer.go:185-255)
func (qa *QueueAnalyzer) Size(t *TargetPerf) (*TargetRate, *AnalysisMetrics, *TargetPerf, error) {
	ctx, err := native.Default()
	if err != nil {
		return nil, nil, nil, err
	}
	rates, m, ach, st, err := ctx.QueueSize([]native.QueueConfig{qa.cfg}, []float32{t.TargetTTFT, t.TargetITL, t.TargetTPS})
	if err != nil {
		return nil, nil, nil, err
	}
	if st[0] != 0 {
		return nil, nil, nil, fmt.Errorf("failed to size queue for targets {TTFT=%.3f, ITL=%.3f, TPS=%.3f}", t.TargetTTFT, t.TargetITL, t.TargetTPS)
	}
	return &TargetRate{rates[0], rates[1], rates[2]}, metricsOf(m[0]), &TargetPerf{ach[0], ach[1], ach[2]}, nil
}
