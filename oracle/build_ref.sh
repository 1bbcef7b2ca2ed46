#!/bin/bash
# Stage the UNMODIFIED reference implementation of the hot path (SURVEY.md §8a files, pure Python + numpy) under
# oracle/_ref/ so that bench.py can time the reference's own fly() on the GPU box's host cores
# (`bench.py --impl reference`, `cpu_baseline.kind = "reference"`).  /root/reference does not exist on the GPU box;
# oracle/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it travels with
# the snapshot like the built .so files.  Files are copied byte for byte — nothing is edited; the two numpy-2 shims
# the reference needs (np.Infinity, scratch cwd) are applied by the caller (oracle/ref_runner.py).
#
#   oracle/build_ref.sh [/root/reference]
set -euo pipefail
REF="${1:-${LD_REFERENCE:-/root/reference}}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/flight" ] || { echo "reference tree not found at $REF" >&2; exit 3; }
rm -rf "$OUT"
mkdir -p "$OUT"
FILES=(
  flight/__init__.py                      # fly()                                   a1, a3
  objects/__init__.py objects/coins.py    # Coin                                    a2
  drone/__init__.py drone/drone.py        # Drone.measure, reward rules             a4, a10
  drone/onboard_computer.py               # ControlUnit ABC (+ PID cascade)         a12
  drone/quadcopter/__init__.py drone/quadcopter/parameters.py
  drone/quadcopter/physical_model.py      # calculate_accelerations                 a5
  drone/quadcopter/geometry.py            # rotation / translational matrices       a6
  drone/quadcopter/choose_paramset.py     # PID gain sets
  drone/sensor/__init__.py drone/sensor/sensor.py   # velocity_verlet               a7, a8
  drone/controller/__init__.py drone/controller/pid.py   #                          a11
  evolve/__init__.py evolve/brain.py evolve/util.py      #                          a9, f-1
  datahandling/__init__.py datahandling/data_scraper.py  # constructed by Drone.__init__ (never used when visualize=False)
)
for f in "${FILES[@]}"; do
  case "$f" in \#*) continue;; esac
  mkdir -p "$OUT/$(dirname "$f")"
  cp "$REF/$f" "$OUT/$f"
done
( cd "$REF" && sha256sum "${FILES[@]}" ) > "$OUT/SHA256SUMS"
echo "staged $(wc -l < "$OUT/SHA256SUMS") reference files under $OUT"
