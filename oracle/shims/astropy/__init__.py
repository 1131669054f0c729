"""Import shim (oracle scaffolding only): names only; grid_tools are never exercised."""
