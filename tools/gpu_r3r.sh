#!/bin/bash
# last library build of the round: VQT / CQT / config parity tests and smoke
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" 2>&1 | tail -2 | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
