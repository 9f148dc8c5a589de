"""Run one of the reference's scripts UNCHANGED on the CUDA path:

    python -m python_5gtoolbox_b200.run_reference_script scripts.mixed_MS_ldpc_search_best_pair --ref /path/to/python_5gtoolbox

The reference's scripts use relative paths (tables, default configuration JSON, out/) and write into out/
(scripts/sim_ldpc_decoder.py:39, scripts/internal/default_config_files.py:7), so they have to run from a directory laid
out like the reference root; this builds a scratch directory of symlinks with its own writable out/, puts it on sys.path,
calls python_5gtoolbox_b200.install() -- which rebinds the LDPC functions inside the reference's own modules and stubs
matplotlib when the box has none -- and runs the module as __main__.  Not a byte of the reference is edited."""
import argparse
import os
import runpy
import sys
import tempfile
import time


def make_workdir(ref, workdir=None):
    d = workdir or tempfile.mkdtemp(prefix="py5g_b200_run_")
    os.makedirs(d, exist_ok=True)
    for name in os.listdir(ref):
        if name == "out" or os.path.lexists(os.path.join(d, name)):
            continue
        os.symlink(os.path.join(os.path.abspath(ref), name), os.path.join(d, name))
    out = os.path.join(d, "out")
    os.makedirs(out, exist_ok=True)
    ref_out = os.path.join(ref, "out")
    if os.path.isdir(ref_out):   # the scripts shipped with sim_flag = 0 only re-read the reference's stored results
        import shutil
        for name in os.listdir(ref_out):
            if not os.path.exists(os.path.join(out, name)) and os.path.isfile(os.path.join(ref_out, name)):
                shutil.copy(os.path.join(ref_out, name), os.path.join(out, name))
    return d


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("module", help="e.g. scripts.mixed_MS_ldpc_search_best_pair")
    ap.add_argument("--ref", required=True, help="root of an unmodified xu753x/python_5gtoolbox checkout")
    ap.add_argument("--workdir", help="scratch directory (default: a new temporary one); results land in its out/")
    ap.add_argument("--seed", type=int, help="np.random.seed before the script starts (the scripts themselves never seed)")
    args = ap.parse_args(argv)
    wd = make_workdir(args.ref, args.workdir)
    os.chdir(wd)
    sys.path.insert(0, wd)
    import python_5gtoolbox_b200
    names = python_5gtoolbox_b200.install()
    print(f"[run_reference_script] {len(names)} names rebound, cwd={wd}", file=sys.stderr)
    if args.seed is not None:
        import numpy as np
        np.random.seed(args.seed)
    t0 = time.time()
    runpy.run_module(args.module, run_name="__main__")
    print(f"[run_reference_script] {args.module} finished in {time.time() - t0:.1f} s; outputs in {os.path.join(wd, 'out')}: "
          f"{sorted(os.listdir(os.path.join(wd, 'out')))}", file=sys.stderr)


if __name__ == "__main__":
    main()
