"""Host-side helpers with the reference's module names: heuristics, wrappers, demands, graph."""
