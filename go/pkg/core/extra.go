// extra.go — the remainder of the reference's exported pkg/core surface (ADVICE r01): Allocation.Scale /
// ReAllocate / TransitionPenalty, AllocationDiff, the Remove* mutators, GetCandidateAccelerators, the
// ServiceClass / Model accessors and the String methods.  WRITTEN WITHOUT A GO TOOLCHAIN; never compiled.
package core

import (
	"bytes"
	"fmt"

	"github.com/llm-d-incubation/workload-variant-autoscaler/pkg/config"
)

// TransitionPenalty (allocation.go:291-300): float32, same operation order.
func (a *Allocation) TransitionPenalty(b *Allocation) float32 {
	if a.accelerator == b.accelerator {
		if a.numReplicas == b.numReplicas {
			return 0
		}
		return b.cost - a.cost
	}
	return config.AccelPenaltyFactor*(a.cost+b.cost) + (b.cost - a.cost)
}

// Scale (allocation.go:165-190): CreateAllocation on the allocation's own accelerator.  CreateAllocation here
// ignores keepAccelerator (it is a Server.Calculate filter, server.go:70-82), as in the reference.
func (a *Allocation) Scale(serverName string) (alloc *Allocation, inc int) {
	server := GetServer(serverName)
	if server == nil || server.Load() == nil || GetAccelerator(a.accelerator) == nil {
		return nil, 0
	}
	alloc = CreateAllocation(serverName, a.accelerator)
	if alloc == nil { // the reference dereferences nil here (allocation.go:188); report "no change" instead of panicking
		return nil, 0
	}
	return alloc, alloc.numReplicas - a.numReplicas
}

// ReAllocate (allocation.go:192-207): cheapest CreateAllocation over all accelerators (minVal == 0 quirk kept);
// accelerators visited in ascending name order (the reference's map order is random).
func (a *Allocation) ReAllocate(serverName string) (*Allocation, string) {
	minVal := float32(0)
	var minAlloc *Allocation
	for _, gName := range sortedKeys(GetAccelerators()) {
		if alloc := CreateAllocation(serverName, gName); alloc != nil {
			if minVal == 0 || alloc.value < minVal {
				minVal = alloc.value
				minAlloc = alloc
			}
		}
	}
	if minAlloc == nil {
		return nil, ""
	}
	return minAlloc, minAlloc.accelerator
}

// AllocationDiff (allocation.go:338-379)
type AllocationDiff struct {
	oldAccelerator string
	newAccelerator string
	oldNumReplicas int
	newNumReplicas int
	costDiff       float32
}

func CreateAllocationDiff(a *Allocation, b *Allocation) *AllocationDiff {
	if a == nil && b == nil {
		return nil
	}
	d := &AllocationDiff{oldAccelerator: "none", newAccelerator: "none"}
	var oldCost, newCost float32
	if a != nil {
		d.oldAccelerator, d.oldNumReplicas, oldCost = a.accelerator, a.numReplicas, a.cost
	}
	if b != nil {
		d.newAccelerator, d.newNumReplicas, newCost = b.accelerator, b.numReplicas, b.cost
	}
	d.costDiff = newCost - oldCost
	return d
}
func (d *AllocationDiff) String() string {
	return fmt.Sprintf("{ %s -> %s, %d -> %d, %v }", d.oldAccelerator, d.newAccelerator, d.oldNumReplicas, d.newNumReplicas, d.costDiff)
}

// ---- System mutators (system.go:104-111, 144-151, 186-193, 249-256) ---------------------------------------
func (s *System) RemoveAccelerator(name string) error {
	if s.accelerators[name] == nil {
		return fmt.Errorf("accelerator %s not found", name)
	}
	delete(s.accelerators, name)
	s.uploaded = false
	return nil
}
func (s *System) RemoveModel(name string) error {
	if s.models[name] == nil {
		return fmt.Errorf("model %s not found", name)
	}
	delete(s.models, name)
	s.uploaded = false
	return nil
}
func (s *System) RemoveServiceClass(name string) error {
	if s.serviceClasses[name] == nil {
		return fmt.Errorf("service class %s not found", name)
	}
	delete(s.serviceClasses, name)
	s.uploaded = false
	return nil
}
func (s *System) RemoveCapacity(name string) bool {
	if _, ok := s.capacity[name]; !ok {
		return false
	}
	delete(s.capacity, name)
	s.uploaded = false
	return true
}
func (s *System) SetAcceleratorsFromSpec(d *config.AcceleratorData) {
	for i := range d.Spec {
		s.AddAcceleratorFromSpec(d.Spec[i])
	}
}
func (s *System) SetCapacityFromSpec(d *config.CapacityData) {
	for _, c := range d.Count {
		s.SetCountFromSpec(c)
	}
}
func (s *System) SetModelsFromSpec(d *config.ModelData) {
	for i := range d.PerfData {
		pd := &d.PerfData[i]
		m := s.models[pd.Name]
		if m == nil {
			m = s.AddModel(pd.Name)
		}
		m.AddPerfDataFromSpec(pd)
	}
	s.uploaded = false
}
func (s *System) SetServersFromSpec(d *config.ServerData) {
	for i := range d.Spec {
		s.AddServerFromSpec(d.Spec[i])
	}
}
func (s *System) SetServiceClassesFromSpec(d *config.ServiceClassData) {
	for i := range d.Spec {
		s.serviceClasses[d.Spec[i].Name] = NewServiceClassFromSpec(&d.Spec[i])
	}
	s.uploaded = false
}
func (s *System) String() string {
	var b bytes.Buffer
	fmt.Fprintf(&b, "Solution: \n")
	for _, n := range sortedKeys(s.servers) {
		fmt.Fprintf(&b, "%s\n", s.servers[n])
	}
	fmt.Fprintf(&b, "AllocationByType: \n")
	for _, n := range sortedKeys(s.allocationByType) {
		fmt.Fprintf(&b, "%s\n", s.allocationByType[n])
	}
	return b.String()
}

// ---- Server (server.go:70-82, 163-170) ---------------------------------------------------------------------
func (s *Server) GetCandidateAccelerators(accelerators map[string]*Accelerator) map[string]*Accelerator {
	if s.keepAccelerator && s.curAllocation != nil {
		if cur := s.curAllocation.accelerator; cur != "" {
			out := map[string]*Accelerator{}
			if g, ok := accelerators[cur]; ok {
				out[cur] = g
			}
			return out
		}
	}
	return accelerators
}
func (s *Server) String() string {
	return fmt.Sprintf("Server: name=%s; class=%s; model=%s; load=%v; allocation=%v", s.name, s.serviceClassName, s.modelName, s.load, s.allocation)
}

// ---- ServiceClass (serviceclass.go:60-103) -----------------------------------------------------------------
func (c *ServiceClass) UpdateModelTargets(spec *config.ServiceClassSpec) bool {
	if spec.Name != c.name || spec.Priority != c.priority {
		return false
	}
	for i := range spec.ModelTargets {
		c.AddModelTarget(&spec.ModelTargets[i])
	}
	return true
}
func (c *ServiceClass) RemoveModelTarget(modelName string) { delete(c.targets, modelName) }
func (c *ServiceClass) Spec() config.ServiceClassSpec {
	mts := make([]config.ModelTarget, 0, len(c.targets))
	for _, m := range sortedKeys(c.targets) {
		t := c.targets[m]
		mts = append(mts, config.ModelTarget{Model: m, SLO_ITL: t.ITL, SLO_TTFT: t.TTFT, SLO_TPS: t.TPS})
	}
	return config.ServiceClassSpec{Name: c.name, Priority: c.priority, ModelTargets: mts}
}
func (c *ServiceClass) String() string {
	return fmt.Sprintf("ServiceClass: name=%s; priority=%d; targets=%v", c.name, c.priority, c.targets)
}
func (t *Target) String() string { return fmt.Sprintf("[ITL=%v, TTFT=%v, TPS=%v]", t.ITL, t.TTFT, t.TPS) }

// ---- Model (model.go:56-75) ----------------------------------------------------------------------------------
func (m *Model) RemovePerfData(accName string) { delete(m.perfData, accName); delete(m.numInstances, accName) }
func (m *Model) Spec() *config.ModelData {
	md := &config.ModelData{PerfData: make([]config.ModelAcceleratorPerfData, 0, len(m.perfData))}
	for _, a := range sortedKeys(m.perfData) {
		md.PerfData = append(md.PerfData, *m.perfData[a])
	}
	return md
}
