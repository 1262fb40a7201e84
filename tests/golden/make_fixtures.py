"""Regenerates tests/golden/*.json from the reference's own fixture files.

Run in the authoring container only (needs /root/reference).  The outputs are the
policy / routing DOCUMENTS of BASELINE configs 1 and 5, derived by the same merge
steps the reference performs at load time:
  policy  = config/safety.yaml  (+) examples/<pack>/overlays/policy.fragment.yaml   (kernel.go:694-711)
  routing = config/pools.yaml   (+) examples/<pack>/overlays/pools.patch.yaml       (json_merge_patch, pack.yaml)
plus the expected decision named by the pack's own policy simulation
(examples/hello-pack/pack.yaml:43-50).
"""
import json
import os
import sys

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cordum_b200 import policy_io  # noqa: E402

REF = "/root/reference"


def read(p):
    with open(os.path.join(REF, p)) as f:
        return f.read()


def build(pack):
    base = policy_io.parse_safety_policy(read("config/safety.yaml"))
    frag = policy_io.parse_safety_policy(read("examples/%s/overlays/policy.fragment.yaml" % pack))
    policy = policy_io.merge_policies(base, frag)
    pools = yaml.safe_load(read("config/pools.yaml"))
    patch = yaml.safe_load(read("examples/%s/overlays/pools.patch.yaml" % pack))
    routing = policy_io.parse_pools_config(policy_io.json_merge_patch(pools, patch))
    return {"policy": policy, "routing": routing}


def main():
    c1 = build("hello-pack")
    pack = yaml.safe_load(read("examples/hello-pack/pack.yaml"))
    c1["policy_simulations"] = pack["tests"]["policySimulations"]
    c5 = build("demo-guardrails")
    for name, doc in (("c1_hello_pack.json", c1), ("c5_demo_guardrails.json", c5)):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(doc, f, indent=1, sort_keys=True)
            f.write("\n")
        print("wrote", name)


if __name__ == "__main__":
    main()
