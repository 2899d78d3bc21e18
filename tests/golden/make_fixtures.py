#!/usr/bin/env python3
"""Generate the committed geography fixture from the read-only reference checkout.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_fixtures.py

Outputs (committed):
  tests/golden/geography.npz       columnar base tables (1 universe)
  tests/golden/geography_meta.json row counts + sha256 of every array, for drift detection

What it restates (nothing is copied; both inputs are parsed where they lie):
  * zips.jsonl                                  -> zips / cities tables
    parse rules follow geography-loader/.../GeographiesLoader.java:51-85:
      `_id` -> Integer.parseInt (:62, leading zeros dropped), `pop` -> int (:63),
      city identity = (city, state) (:69-71; geography/.../City.java:7), first-seen city
      fixes its state (:82-84), `loc` ignored.
  * geography-loader/.../StateData.java:21-72   -> states table (52 entries, AL twice => 51 rows)
  * geography-loader/.../StateData.java:78-296  -> 219 directed adjacency pairs, appended per
    state in list order exactly like Association.add does in app/.../Runner.java:172-193.

Row order.  The Java app iterates HashSets (Runner.java:97,124,150) so its row order is a JVM
implementation detail.  The engine is order-agnostic; our canonical order is: ZIPs in file order,
cities in first-appearance order, states in STATES list order de-duplicated.
"""
from __future__ import annotations

import hashlib
import json
import re
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT_DIR = Path(__file__).resolve().parent


def parse_state_data(java_src: str):
    states = re.findall(r'new State\("([^"]+)",\s*"([^"]+)"\)', java_src)
    adj = re.findall(r'new StateAdjacency\("([^"]+)",\s*"([^"]+)"\)', java_src)
    seen, uniq = set(), []
    for code, name in states:  # HashSet<State>(STATES) de-duplicates the record (code, name)
        if (code, name) not in seen:
            seen.add((code, name))
            uniq.append((code, name))
    return states, uniq, adj


def pack_strings(strings):
    offsets = np.zeros(len(strings) + 1, dtype=np.uint32)
    blobs = [s.encode("utf-8") for s in strings]
    np.cumsum([len(b) for b in blobs], out=offsets[1:])
    data = np.frombuffer(b"".join(blobs), dtype=np.uint8).copy()
    return offsets, data


def main() -> int:
    if not REF.exists():
        print("reference checkout not present; fixtures can only be regenerated in the build container",
              file=sys.stderr)
        return 1

    raw_states, states, adj_pairs = parse_state_data(
        (REF / "geography-loader/src/main/java/dgroomes/geography_loader/StateData.java").read_text())
    assert len(raw_states) == 52 and len(states) == 51 and len(adj_pairs) == 219, \
        (len(raw_states), len(states), len(adj_pairs))
    state_index = {code: i for i, (code, _n) in enumerate(states)}

    # adjacency as per-state target lists in list order (Runner.java:172-193)
    adj_lists = [[] for _ in states]
    for a, b in adj_pairs:
        adj_lists[state_index[a]].append(state_index[b])
    adj_offsets = np.zeros(len(states) + 1, dtype=np.int32)
    np.cumsum([len(l) for l in adj_lists], out=adj_offsets[1:])
    adj_targets = np.array([t for l in adj_lists for t in l], dtype=np.int32)

    zip_code, zip_pop, zip_city = [], [], []
    city_index: dict[tuple[str, str], int] = {}
    city_names, city_state = [], []
    seen_zip = set()
    with open(REF / "zips.jsonl", "r", encoding="utf-8") as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            node = json.loads(line)
            code = int(node["_id"])          # Integer.parseInt("01001") == 1001
            pop = int(node["pop"])
            key = (node["city"], node["state"])
            if (code, pop) in seen_zip:      # Set<Zip> semantics (record equality)
                continue
            seen_zip.add((code, pop))
            if key not in city_index:
                city_index[key] = len(city_names)
                city_names.append(node["city"])
                city_state.append(state_index[node["state"]])
            zip_code.append(code)
            zip_pop.append(pop)
            zip_city.append(city_index[key])

    # app/src/test/java/dgroomes/TheTest.java:22-26
    assert len(zip_code) == 29_353 and len(city_names) == 25_701 and len(states) == 51

    city_off, city_bytes = pack_strings(city_names)
    code_off, code_bytes = pack_strings([c for c, _ in states])
    name_off, name_bytes = pack_strings([n for _, n in states])

    arrays = dict(
        zip_code=np.array(zip_code, dtype=np.int32),
        zip_pop=np.array(zip_pop, dtype=np.int32),
        zip_city=np.array(zip_city, dtype=np.int32),
        city_name_offsets=city_off,
        city_name_bytes=city_bytes,
        city_state=np.array(city_state, dtype=np.int32),
        state_code_offsets=code_off,
        state_code_bytes=code_bytes,
        state_name_offsets=name_off,
        state_name_bytes=name_bytes,
        adj_offsets=adj_offsets,
        adj_targets=adj_targets,
    )
    np.savez_compressed(OUT_DIR / "geography.npz", **arrays)
    meta = {
        "source": "zips.jsonl + StateData.java of dgroomes/java-columnar-query-engine (parsed, not copied)",
        "rows": {"zips": len(zip_code), "cities": len(city_names), "states": len(states),
                 "adjacency_edges": int(adj_targets.size), "city_name_bytes": int(city_bytes.size)},
        "sha256": {k: hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() for k, v in arrays.items()},
    }
    (OUT_DIR / "geography_meta.json").write_text(json.dumps(meta, indent=1, sort_keys=True) + "\n")
    print(json.dumps(meta["rows"]))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
