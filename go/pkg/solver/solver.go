// Package solver — drop-in replacement of the reference's pkg/solver (optimizer.go, solver.go,
// greedy.go): same exported identifiers; SolveUnlimited / SolveGreedy run in the B200 library.
// WRITTEN WITHOUT A GO TOOLCHAIN; never compiled (INTEGRATION.md).
package solver

import (
	"fmt"

	"github.com/llm-d-incubation/workload-variant-autoscaler/go/pkg/core"
	"github.com/llm-d-incubation/workload-variant-autoscaler/pkg/config"
)

type Solver struct {
	optimizerSpec     *config.OptimizerSpec
	currentAllocation map[string]*core.Allocation
	diffAllocation    map[string]*AllocationDiff
	solutionTimeUsec  int64
}

// AllocationDiff mirrors core.AllocationDiff of the reference (allocation.go:338-379).
type AllocationDiff struct {
	OldAccelerator, NewAccelerator string
	OldNumReplicas, NewNumReplicas int
	CostDiff                       float32
}

func NewSolver(spec *config.OptimizerSpec) *Solver {
	return &Solver{optimizerSpec: spec, currentAllocation: map[string]*core.Allocation{}, diffAllocation: map[string]*AllocationDiff{}}
}

// Solve (solver.go:32-60)
func (s *Solver) Solve() error {
	s.currentAllocation = map[string]*core.Allocation{}
	for name, server := range core.GetServers() {
		if a := server.CurAllocation(); a != nil {
			s.currentAllocation[name] = a
		}
	}
	usec, err := core.TheSystem.Solve(s.optimizerSpec)
	if err != nil {
		return err
	}
	s.solutionTimeUsec = usec
	s.diffAllocation = map[string]*AllocationDiff{}
	for name, server := range core.GetServers() {
		cur, des := s.currentAllocation[name], server.Allocation()
		if cur == nil && des == nil {
			continue
		}
		d := &AllocationDiff{OldAccelerator: "none", NewAccelerator: "none"}
		var oldCost, newCost float32
		if cur != nil {
			d.OldAccelerator, d.OldNumReplicas, oldCost = cur.Accelerator(), cur.NumReplicas(), cur.Cost()
		}
		if des != nil {
			d.NewAccelerator, d.NewNumReplicas, newCost = des.Accelerator(), des.NumReplicas(), des.Cost()
		}
		d.CostDiff = newCost - oldCost
		s.diffAllocation[name] = d
	}
	return nil
}

// SolveUnlimited / SolveGreedy keep the reference's entry points (solver.go:63, greedy.go:35).
func (s *Solver) SolveUnlimited() { spec := *s.optimizerSpec; spec.Unlimited = true; _, _ = core.TheSystem.Solve(&spec) }
func (s *Solver) SolveGreedy()    { spec := *s.optimizerSpec; spec.Unlimited = false; _, _ = core.TheSystem.Solve(&spec) }
func (s *Solver) AllocationDiff() map[string]*AllocationDiff { return s.diffAllocation }

type Optimizer struct {
	spec             *config.OptimizerSpec
	solver           *Solver
	solutionTimeMsec int64
}

func NewOptimizerFromSpec(spec *config.OptimizerSpec) *Optimizer { return &Optimizer{spec: spec} }

// Optimize (optimizer.go:24-35); SolutionTimeMsec is the device+host time of the native solve.
func (o *Optimizer) Optimize() error {
	if o.spec == nil {
		return fmt.Errorf("missing optimizer spec")
	}
	o.solver = NewSolver(o.spec)
	err := o.solver.Solve()
	o.solutionTimeMsec = o.solver.solutionTimeUsec / 1000
	return err
}
func (o *Optimizer) SolutionTimeMsec() int64 { return o.solutionTimeMsec }
