set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import sys; sys.path.insert(0,'.')
from phaseshifts_b200 import api
print('fp64 peak TF/s, implied MHz:', api.fp64_peak(0))
PY
