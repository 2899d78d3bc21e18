#!/bin/bash
# per-kernel SASS mnemonic counts of the shipped library (evidence for TMA / vectorised loads / barriers / peer stores)
cd /root/repo
SO=java-columnar-query-engine_b200/lib/libcolq.so
OUT=profiles/r02_sass_counts.txt
{
echo "# cuobjdump -sass $SO (sm_100a) -- per-kernel counts of the instructions that carry the design:"
echo "#   UBLKCP = cp.async.bulk (TMA 1-D bulk copy)   SYNCS = mbarrier ops   LDG.E.128 = 128-bit global loads"
echo "#   ATOMG/RED = global atomics   VOTE = ballots   SHFL = shuffles   ACQBULK/griddepcontrol -> 'ACQBULK'/'PREEXIT' style ops listed if present"
echo "# build: $(python -c "import ctypes;l=ctypes.CDLL('$SO');l.colq_build_id.restype=ctypes.c_char_p;print(l.colq_build_id().decode())")"
printf "%-70s %7s %7s %7s %9s %6s %6s %6s %6s %6s\n" kernel instrs UBLKCP SYNCS LDG.E.128 LDG STG ATOM VOTE SHFL
cuobjdump -sass $SO 2>/dev/null | awk '
/Function :/ { if (name != "") emit(); name=$3; n=ub=sy=l128=ldg=stg=at=vo=sh=0; next }
/^ +\/\*[0-9a-f]+\*\// { n++; if ($0 ~ /UBLKCP/) ub++; if ($0 ~ /SYNCS/) sy++; if ($0 ~ /LDG\.E\.(128|ENL2\.256|.*\.128)/ || $0 ~ /LDG\.E\.128/ || $0 ~ /LDG.*\.128/) l128++; if ($0 ~ / LDG/) ldg++; if ($0 ~ / STG/) stg++; if ($0 ~ /ATOMG|ATOM\.|REDG| RED\./) at++; if ($0 ~ /VOTE/) vo++; if ($0 ~ /SHFL/) sh++ }
function emit() { printf "%-70s %7d %7d %7d %9d %6d %6d %6d %6d %6d\n", name, n, ub, sy, l128, ldg, stg, at, vo, sh }
END { emit() }' | while read -r name rest; do printf "%-70s %s\n" "$(echo $name | c++filt | sed 's/colq:://; s/(.*//' | cut -c1-70)" "$rest"; done | sort
echo
echo "# programmatic dependent launch instructions (griddepcontrol.*):"
cuobjdump -sass $SO 2>/dev/null | grep -E "ACQBULK|PREEXIT|DEPBAR.*LE|griddep" | awk '{print $2}' | sort | uniq -c | head
} > $OUT
wc -l $OUT; head -60 $OUT
