#!/usr/bin/env python
"""tests/golden/simu_data_reference.json <- /root/reference/tests/testthat/fixtures/simu_data_reference.rds

The reference's own golden fixture (exact simu_data() outputs of six configurations and seeds,
tests/testthat/test-simu_data-characterization.R:11-52) decoded with scripts/rds_reader.py, so that tests/test_rsim.py can
check magmaclustr_b200/rsim.py where /root/reference does not exist.  Run in the BUILD container only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from rds_reader import read_rds                     # noqa: E402

FIXTURE = "/root/reference/tests/testthat/fixtures/simu_data_reference.rds"


def plain(o):
    if isinstance(o, dict):
        return {k: plain(v) for k, v in o.items() if not k.startswith("__")}
    if isinstance(o, (list, tuple)):
        return [plain(v) for v in o]
    return o


def main():
    obj = read_rds(FIXTURE)
    names = list(obj["configs"]["__names__"])
    out = {"source": "tests/testthat/fixtures/simu_data_reference.rds (reference-held)", "names": names,
           "seeds": {n: obj["seeds"][i] for i, n in enumerate(names)},
           "configs": {n: plain(obj["configs"][n]) for n in names},
           "columns": {n: list(obj["results"][n]["__names__"]) for n in names},
           "results": {n: plain(obj["results"][n]) for n in names}}
    path = os.path.join(ROOT, "tests", "golden", "simu_data_reference.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
