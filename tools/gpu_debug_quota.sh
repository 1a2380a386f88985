#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/q && cd /tmp/q
python - <<'PY'
import sys, os, subprocess
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
from util import unpack_case
d = unpack_case("synth_small", "/tmp/q")
cli="/root/repo/rnaseqc_b200/bin/rnaseqc"
def run(tag,*flags):
    out=d+"/o_"+tag
    p=subprocess.run([cli, d+"/a.gtf", d+"/a.sam", out, "--coverage","--bed",d+"/a.bed","--fragment-samples","300",*flags],stdout=subprocess.PIPE,stderr=subprocess.PIPE,text=True)
    n=sum(1 for _ in open(out+"/a.sam.fragmentSizes.txt"))
    tot=sum(int(l.split()[1]) for l in open(out+"/a.sam.fragmentSizes.txt") if not l.startswith("Fragment"))
    print(tag, p.returncode, "lines", n, "samples", tot, p.stderr[-200:], flush=True)
run("default")
run("nocompact","--no-compact-submit")
run("plain","--no-packed-submit")
run("smallbatch","--batch-records","1500")
run("smallbatch_plain","--batch-records","1500","--no-packed-submit")
run("seq","--decode-threads","0")
PY
