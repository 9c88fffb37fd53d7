#!/bin/bash
# End-to-end check of the command line on 1 and 2 GPUs: same pairs file, outputs must be byte-identical.
# usage (on a box with >= 2 GPUs): bash tools/cli_multigpu_check.sh [n_pairs] [T]
set -e
N=${1:-96}; T=${2:-1200}
D=$(mktemp -d)
python - "$D" "$N" "$T" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from poreover_b200 import synth
d, n, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
with open(os.path.join(d, "pairs.txt"), "w") as f:
    for k in range(n):
        f1, f2 = synth.save_pair(d, 100 + k, T + 37 * (k % 9))
        f.write("%s %s\n" % (f1, f2))
PY
python -m poreover_b200 pair-decode "$D/pairs.txt" --dir "$D" --basecaller bonito --reverse_complement --beam_width 25 --out "$D/one" 2> "$D/one.err"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 -m poreover_b200 pair-decode "$D/pairs.txt" --dir "$D" --basecaller bonito --reverse_complement --beam_width 25 --out "$D/two" 2> "$D/two.err"
for ext in 1d.fasta 2d.fasta; do cmp "$D/one.$ext" "$D/two.$ext"; done
diff <(grep -v '^# {' "$D/one.log") <(grep -v '^# {' "$D/two.log")
python -m poreover_b200 decode "$D" --basecaller bonito --algorithm beam --beam_width 5 --out "$D/dec_one" 2> "$D/dec_one.err"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 -m poreover_b200 decode "$D" --basecaller bonito --algorithm beam --beam_width 5 --out "$D/dec_two" 2> "$D/dec_two.err"
cmp "$D/dec_one.fasta" "$D/dec_two.fasta"
echo "decode: $(grep -c '>' "$D/dec_one.fasta") records identical on 1 and 2 GPUs"
echo "CLI 1-GPU and 2-GPU outputs identical: $(grep -c '>' "$D/one.2d.fasta") consensus records, $(wc -c < "$D/one.2d.fasta") bytes"
