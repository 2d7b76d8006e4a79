package core

import "sort"

// sortedKeys fixes the dense index order (the reference iterates Go maps in random order; only exact
// ties can observe the difference, DESIGN.md section 2).
func sortedKeys[V any](m map[string]V) []string {
	keys := make([]string, 0, len(m))
	for k := range m {
		keys = append(keys, k)
	}
	sort.Strings(keys)
	return keys
}
