"""Where the time of the patched GMMReg.forward goes (torch.profiler, CUDA kernel table).

    python tools/profile_model.py [pairs] > gpurun_out/model_profile.txt

Needs baseline/_ref (vendored by __graft_entry__.build()).  Prints the kernel table of the unpatched and the patched
forward so that the share of the repo's own kernels against the PyTorch remainder (DGCNN conv2..5, transformer) is on
record.  Profiler numbers are for shares only; bench.py times the forward without a profiler.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    from oracle import refload
    import ogmm_b200.install as inst
    from ogmm_b200 import synth
    ref = refload.import_reference()
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    model = ref["gmmreg"].GMMReg(512, 16, refload.model_config()).to(dev).eval()
    s, t, _, _ = synth.modelnet_batch(0, pairs, 1024)
    src, tgt = torch.from_numpy(s).to(dev), torch.from_numpy(t).to(dev)

    def run(tag):
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=False):
            for _ in range(2):
                torch.manual_seed(7)
                model(src, tgt)
            torch.cuda.synchronize()
            with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
                torch.manual_seed(7)
                model(src, tgt)
                torch.cuda.synchronize()
        ev = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
        total = sum(e.device_time_total for e in ev)
        print(f"== {tag}: {pairs} pairs, CUDA kernel time {total / 1e3:.2f} ms, {sum(e.count for e in ev)} launches")
        ours = sum(e.device_time_total for e in ev if "ogmm" in e.key)
        print(f"   repo kernels (namespace ogmm): {ours / 1e3:.2f} ms = {100 * ours / max(total, 1):.1f} %")
        for e in sorted(ev, key=lambda e: -e.device_time_total)[:22]:
            print(f"   {e.device_time_total / 1e3:9.3f} ms  x{e.count:<5d} {e.key[:110]}")

    run("reference (stock PyTorch CUDA path)")
    inst.install(model=model)
    try:
        run("after install()")
    finally:
        inst.uninstall()


if __name__ == "__main__":
    main()
