"""Developer tool (run HERE, no GPU needed): prebuild library variants with extra -D flags into
bipartite-link-prediction_b200/variants/ so that one gpurun call can A/B them (built .so files
travel with the snapshot; nvcc on the GPU box would burn box minutes).
usage: build_variants.py name1="-DX=1 -DY=2" name2="" ..."""
import importlib, os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module('bipartite-link-prediction_b200._lib')
vdir = os.path.join(ROOT, 'bipartite-link-prediction_b200', 'variants')
os.makedirs(vdir, exist_ok=True)


def one(arg):
    name, _, flags = arg.partition('=')
    out = os.path.join(vdir, 'libblp_%s.so' % name)
    r = subprocess.run(L.nvcc_command(out=out, extra=tuple(flags.split())), stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    return name, r.returncode, r.stdout[-2000:]


with ThreadPoolExecutor(4) as ex:
    for name, rc, log in ex.map(one, sys.argv[1:]):
        print(name, 'ok' if rc == 0 else 'FAILED\n' + log, flush=True)
