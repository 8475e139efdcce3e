"""TEST INFRASTRUCTURE (build container only).  Writes tests/golden/state_dict_manifest.json: the key
names, shapes and dtypes of `NeRFNetwork.state_dict()` of the reference's OWN model class
(nvsf/nerf/models/network_dynamic.py, imported by file path through oracle/ref_import.py with the
tinycudann stand-in), i.e. what `Trainer.save_checkpoint` stores under "model"
(nvsf/nerf/utils.py:610-650).  tests/test_checkpoint.py checks that
`NeRFNetwork.load_reference_state_dict` consumes exactly these keys.
Usage:  python -m oracle.make_golden_state_dict"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

S = importlib.import_module("selfsupervised-nvsf_b200.synth")


def main():
    nd = ref_import.import_reference()
    model = nd.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                           min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)
    sd = model.state_dict()
    man = {"constructor": dict(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND),
           "optimizer_groups": [],
           "keys": [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()]}
    # the optimiser's parameter groups (network_dynamic.py:335-357): names resolved through id()
    names = {id(p): n for n, p in model.named_parameters()}
    for g in model.get_params(1.0):
        ps = g["params"] if isinstance(g["params"], (list, tuple)) else list(g["params"])
        man["optimizer_groups"].append({"lr": g["lr"], "params": [names[id(p)] for p in ps]})
    out = os.path.join(ROOT, "tests", "golden", "state_dict_manifest.json")
    json.dump(man, open(out, "w"), indent=0)
    print("wrote", out, len(man["keys"]), "keys")


if __name__ == "__main__":
    main()
