// Package manager — drop-in replacement of the reference's pkg/manager/manager.go.
// WRITTEN WITHOUT A GO TOOLCHAIN; never compiled (INTEGRATION.md).
package manager

import (
	"github.com/llm-d-incubation/workload-variant-autoscaler/go/pkg/core"
	"github.com/llm-d-incubation/workload-variant-autoscaler/go/pkg/solver"
)

type Manager struct {
	system    *core.System
	optimizer *solver.Optimizer
}

func NewManager(system *core.System, optimizer *solver.Optimizer) *Manager {
	core.TheSystem = system
	return &Manager{system: system, optimizer: optimizer}
}

// Optimize (manager.go:21-27)
func (m *Manager) Optimize() error {
	if err := m.optimizer.Optimize(); err != nil {
		return err
	}
	m.system.AllocateByType()
	return nil
}
