"""One computeStep inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off`:

    ncu --set full --clock-control none --profile-from-start off -k regex:'<kernels>' -o gpurun_out/x \
        python profiles/capture_step.py [--heat] [--rows R --cols C --soil-layers L] [--warmup W]

Sets up the C2 (or, with --heat, C3) workload of bench.py, runs W untimed steps, then profiles exactly one."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
from criteria3d_b200 import load_product  # noqa: E402
from criteria3d_b200.synth import Catchment, setup  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--heat", action="store_true")
ap.add_argument("--rows", type=int, default=1024)
ap.add_argument("--cols", type=int, default=1024)
ap.add_argument("--soil-layers", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--saturated-bottom", action="store_true")
a = ap.parse_args()
sf = load_product()
cat = Catchment(a.rows, a.cols, a.soil_layers, heat=a.heat, saturated_bottom=a.saturated_bottom)
setup(sf, cat)
assert sf.set_forcing_rasters(precipitation=cat.rain_raster(40.0)) == 0
for _ in range(a.warmup):
    sf.computeStep(3600.0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
dt = sf.computeStep(3600.0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step, dt =", dt, sf.counters())
