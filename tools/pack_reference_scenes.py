"""Pack the reference's benchmark scenes into iactrace_b200/data/*.npz.

Run in the build container (needs /root/reference): reads configs/HESS/CT3.yaml and CT5.yaml,
stores the parsed numbers losslessly, and checks that unpacking reproduces the parsed YAML exactly.
"""
import sys
from pathlib import Path

import yaml

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from iactrace_b200.io.scene_pack import pack_config, unpack_config  # noqa: E402

REF = Path("/root/reference/configs/HESS")
OUT = Path(__file__).resolve().parent.parent / "iactrace_b200" / "data"

for name in ("CT3", "CT5"):
    cfg = yaml.load(open(REF / f"{name}.yaml"), Loader=yaml.CSafeLoader)
    pack_config(cfg, OUT / f"{name}.npz")
    back = unpack_config(OUT / f"{name}.npz")
    assert back == cfg, f"{name}: round trip differs"
    print(name, (OUT / f"{name}.npz").stat().st_size, "bytes, round trip exact")
