#!/bin/bash
# the whole GPU suite and smoke() on the final tree
set -u
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
