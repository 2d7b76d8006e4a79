// Package analyzer — the public surface of the reference's pkg/analyzer over the B200 library, with the
// reference's types, signatures and error behaviour:
//   queueanalyzer.go  NewQueueAnalyzer / BuildModel / Analyze / Size, PrefillTime / DecodeTime,
//                     EvalTTFT / EvalITL, EffectiveConcurrency
//   utils.go          WithinTolerance, BinarySearch, Model, EvalServTime, EvalWaitingTime
//   queuemodel.go, mm1modelstatedependent.go, mm1kmodel.go   the model types and their getters
// Chains run on the device (wva_queue_analyze / wva_queue_size / wva_model_solve); what takes a Go closure
// (BinarySearch) or is two float32 operations (PrefillTime ...) is host Go -- float32 arithmetic in Go is IEEE,
// one rounding per operation, exactly what the device code reproduces.
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

// BuildModel (queueanalyzer.go:99-131): kept for callers that skip the checks of NewQueueAnalyzer.
func BuildModel(c *Configuration, rq *RequestSize) *QueueAnalyzer {
	qa, _ := NewQueueAnalyzer(c, rq)
	return qa
}

// PrefillTime (queueanalyzer.go:257-262) and DecodeTime (:264-266): float32, left-associated.
func (p *PrefillParms) PrefillTime(avgInputTokens int, batchSize float32) float32 {
	if avgInputTokens == 0 {
		return 0
	}
	return p.Gamma + p.Delta*float32(avgInputTokens)*batchSize
}
func (p *DecodeParms) DecodeTime(batchSize float32) float32 { return p.Alpha + p.Beta*batchSize }

// EffectiveConcurrency (queueanalyzer.go:296-302); Go's builtin min/max propagate NaN.
func EffectiveConcurrency(avgServiceTime float32, sp *ServiceParms, rq *RequestSize, maxBatchSize int) float32 {
	tokens := float32(rq.AvgOutputTokens - 1)
	numerator := avgServiceTime - (sp.Prefill.Gamma + sp.Decode.Alpha*tokens)
	denominator := (sp.Prefill.Delta * float32(rq.AvgInputTokens)) + (sp.Decode.Beta * tokens)
	n := numerator / denominator
	return min(max(n, 0), float32(maxBatchSize))
}

// WithinTolerance (utils.go:12-20)
func WithinTolerance(x, value, tolerance float32) bool {
	if x == value {
		return true
	}
	if value == 0 || tolerance < 0 {
		return false
	}
	d := float64((x - value) / value)
	if d < 0 {
		d = -d
	}
	return d <= float64(tolerance)
}

var epsilon float32 = 1e-6
var maxIterations int = 100

// BinarySearch (utils.go:26-70): bisection on a monotone eval; indicator -1 / 0 / +1 as in the reference.
func BinarySearch(xMin float32, xMax float32, yTarget float32, eval func(float32) (float32, error)) (float32, int, error) {
	if xMin > xMax {
		return 0, 0, fmt.Errorf("invalid range [%v, %v]", xMin, xMax)
	}
	yBounds := make([]float32, 2)
	var err error
	for i, x := range []float32{xMin, xMax} {
		if yBounds[i], err = eval(x); err != nil {
			return 0, 0, fmt.Errorf("invalid function evaluation: %v", err)
		}
		if WithinTolerance(yBounds[i], yTarget, epsilon) {
			return x, 0, nil
		}
	}
	increasing := yBounds[0] < yBounds[1]
	if increasing && yTarget < yBounds[0] || !increasing && yTarget > yBounds[0] {
		return xMin, -1, nil
	}
	if increasing && yTarget > yBounds[1] || !increasing && yTarget < yBounds[1] {
		return xMax, +1, nil
	}
	var xStar, yStar float32
	for i := 0; i < maxIterations; i++ {
		xStar = 0.5 * (xMin + xMax)
		if yStar, err = eval(xStar); err != nil {
			return 0, 0, fmt.Errorf("invalid function evaluation: %v", err)
		}
		if WithinTolerance(yStar, yTarget, epsilon) {
			break
		}
		if increasing && yTarget < yStar || !increasing && yTarget > yStar {
			xMax = xStar
		} else {
			xMin = xStar
		}
	}
	return xStar, 0, nil
}

// ---- queue models (queuemodel.go, mm1modelstatedependent.go, mm1kmodel.go) ------------------------------

// QueueModel holds the statistics of the last Solve (queuemodel.go:9-25).
type QueueModel struct {
	lambda, mu, rho                                                     float32
	avgRespTime, avgWaitTime, avgServTime, avgNumInSystem, avgQueueLength float32
	isValid                                                             bool
}

func (m *QueueModel) IsValid() bool              { return m.isValid }
func (m *QueueModel) GetLambda() float32         { return m.lambda }
func (m *QueueModel) GetMu() float32             { return m.mu }
func (m *QueueModel) GetRho() float32            { return m.rho }
func (m *QueueModel) GetAvgQueueLength() float32 { return m.avgQueueLength }
func (m *QueueModel) GetAvgNumInSystem() float32 { return m.avgNumInSystem }
func (m *QueueModel) GetAvgWaitTime() float32    { return m.avgWaitTime }
func (m *QueueModel) GetAvgServTime() float32    { return m.avgServTime }
func (m *QueueModel) GetAvgRespTime() float32    { return m.avgRespTime }

// MM1ModelStateDependent (mm1modelstatedependent.go:9-26).  The probability vector lives on the device side
// of wva_model_solve; the model keeps the history of Solve calls so that the stale-p[0] validity rule of
// queuemodel.go:30 is reproduced: every Solve replays the sequence on a fresh native model (sequences are
// short -- a bisection -- and the native call runs them in one launch).
type MM1ModelStateDependent struct {
	QueueModel
	K               int
	servRate        []float32
	p               []float64
	avgNumInServers float32
	throughput      float32
	lambdas, mus    []float32
}

func NewMM1ModelStateDependent(K int, servRate []float32) *MM1ModelStateDependent {
	return &MM1ModelStateDependent{K: K, servRate: servRate, p: make([]float64, K+1)}
}

// Solve (queuemodel.go:27-37 + mm1modelstatedependent.go:38-116)
func (m *MM1ModelStateDependent) Solve(lambda float32, mu float32) {
	m.lambdas, m.mus = append(m.lambdas, lambda), append(m.mus, mu)
	ctx, err := native.Default()
	if err != nil {
		m.isValid = false
		return
	}
	out, p, err := ctx.ModelSolve(m.K, m.servRate, m.lambdas, m.mus)
	if err != nil {
		m.isValid = false
		return
	}
	o := out[9*(len(m.lambdas)-1):]
	m.lambda, m.mu = lambda, mu
	m.isValid, m.rho, m.avgRespTime, m.avgWaitTime, m.avgServTime = o[0] != 0, o[1], o[2], o[3], o[4]
	m.avgNumInSystem, m.avgQueueLength, m.avgNumInServers, m.throughput = o[5], o[6], o[7], o[8]
	m.p = p
}
func (m *MM1ModelStateDependent) ComputeRho() float32           { return 1 - float32(m.p[0]) }
func (m *MM1ModelStateDependent) GetRhoMax() float32            { return float32(m.K) }
func (m *MM1ModelStateDependent) GetAvgNumInServers() float32   { return m.avgNumInServers }
func (m *MM1ModelStateDependent) GetThroughput() float32        { return m.throughput }
func (m *MM1ModelStateDependent) GetProbabilities() []float64   { return m.p }

// MM1KModel (mm1kmodel.go).  DEVIATION, documented: the reference evaluates the M/M/1/K closed form with
// math.Pow (mm1kmodel.go:63,68); this shim runs the same queue as a state-dependent chain with the constant rate
// mu, which the reference's own test pins to agree within 1e-3 (queuemodel_test.go:461-496) and which agrees to
// ~1e-6 relative in practice.  MM1KModel is not on the CreateAllocation path (SURVEY 8c).
type MM1KModel struct {
	MM1ModelStateDependent
}

func NewMM1KModel(K int) *MM1KModel {
	return &MM1KModel{MM1ModelStateDependent: MM1ModelStateDependent{K: K, p: make([]float64, K+1)}}
}
func (m *MM1KModel) Solve(lambda float32, mu float32) {
	m.servRate = []float32{mu}
	m.MM1ModelStateDependent.Solve(lambda, mu)
	if lambda >= 0 && mu > 0 {
		m.rho = lambda / mu // MM1KModel.ComputeRho is lambda/mu (mm1kmodel.go:37-44)
	}
}
func (m *MM1KModel) ComputeRho() float32 {
	if m.mu == 0 {
		return 0
	}
	return m.lambda / m.mu
}

// Model and the Eval* helpers of utils.go:73-92 / queueanalyzer.go:270-290 (package-level state, as in the reference).
var Model *MM1ModelStateDependent
var evalRequestSize *RequestSize
var evalServiceParms *ServiceParms
var evalMaxBatchSize int

func EvalServTime(x float32) (float32, error) {
	Model.Solve(x, 1)
	if !Model.IsValid() {
		return 0, fmt.Errorf("invalid model %v", Model)
	}
	return Model.GetAvgServTime(), nil
}
func EvalWaitingTime(x float32) (float32, error) {
	Model.Solve(x, 1)
	if !Model.IsValid() {
		return 0, fmt.Errorf("invalid model %v", Model)
	}
	return Model.GetAvgWaitTime(), nil
}
func EvalTTFT(x float32) (float32, error) {
	Model.Solve(x, 1)
	if !Model.IsValid() {
		return 0, fmt.Errorf("invalid model %v", Model)
	}
	effConc := EffectiveConcurrency(Model.GetAvgServTime(), evalServiceParms, evalRequestSize, evalMaxBatchSize)
	return Model.GetAvgWaitTime() + evalServiceParms.Prefill.PrefillTime(evalRequestSize.AvgInputTokens, effConc), nil
}
func EvalITL(x float32) (float32, error) {
	Model.Solve(x, 1)
	if !Model.IsValid() {
		return 0, fmt.Errorf("invalid model %v", Model)
	}
	effConc := EffectiveConcurrency(Model.GetAvgServTime(), evalServiceParms, evalRequestSize, evalMaxBatchSize)
	return evalServiceParms.Decode.DecodeTime(effConc), nil
}

// SetEvalContext points Model and the eval globals at this analyzer (what Size does at queueanalyzer.go:196-203
// before its bisections), so that EvalTTFT / EvalITL can be called by reference tests.
func (qa *QueueAnalyzer) SetEvalContext() {
	n := qa.MaxBatchSize
	serv := make([]float32, n)
	for i := 1; i <= n; i++ {
		numDecode := qa.RequestSize.AvgOutputTokens - 1
		if qa.RequestSize.AvgInputTokens == 0 && qa.RequestSize.AvgOutputTokens == 1 {
			numDecode = 1
		}
		pre := qa.ServiceParms.Prefill.PrefillTime(qa.RequestSize.AvgInputTokens, float32(i))
		dec := float32(numDecode) * qa.ServiceParms.Decode.DecodeTime(float32(i))
		serv[i-1] = float32(i) / (pre + dec)
	}
	Model = NewMM1ModelStateDependent(qa.MaxBatchSize+qa.MaxQueueSize, serv)
	evalRequestSize, evalServiceParms, evalMaxBatchSize = qa.RequestSize, qa.ServiceParms, qa.MaxBatchSize
}
