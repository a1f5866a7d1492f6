for c in c0 c64 u64 c128 u32a c192 u16a c320 u8a u8b; do timeout 60 python scripts/prof_conv.py $c ablate 2>&1 | sed -n 2,2p | cut -c1-62; done
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from view_fusion_b200 import _lib
lib=_lib.require_device()
PY
