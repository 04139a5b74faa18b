cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/exp6_debug.log 2>&1
import sys, os
sys.argv = ["conv_sweep"]
sys.path.insert(0, "tools")
import importlib.util, torch
spec = importlib.util.spec_from_file_location("cs", "tools/conv_sweep.py")
src = open("tools/conv_sweep.py").read().replace("\nmain()\n", "\n")
ns = {"__file__": os.path.abspath("tools/conv_sweep.py")}
exec(compile(src, "tools/conv_sweep.py", "exec"), ns)
lib = ns["_lib"].lib()
shapes = [(1024, 112, 112, 27, 64, 1, 1, 2, 0, 0), (1024, 112, 112, 64, 64, 3, 1, 2, 0, 1), (1024, 56, 56, 64, 64, 3, 1, 0, 1, 0),
          (1024, 56, 56, 64, 128, 3, 1, 2, 0, 1), (1024, 28, 28, 128, 128, 3, 1, 0, 1, 0), (1024, 112, 112, 64, 64, 3, 2, 0, 1, 0),
          (64, 320, 320, 28, 28, 3, 1, 1, 0, 0), (64, 160, 160, 56, 56, 3, 1, 1, 1, 0), (64, 80, 80, 88, 88, 3, 1, 1, 1, 0)]
flags = [0, 1, 16, 17, 4, 8, 12, 20, 28]
print("shape".ljust(44) + "".join(f"D{f}".rjust(9) for f in flags))
for sh in shapes:
    row = []
    for f in flags:
        for k, v in ns["DEFAULTS"].items():
            ns["_lib"].check(lib.b2f_set_tuning(k, v))
        ns["_lib"].check(lib.b2f_set_tuning(4, f))
        ms, tf, gb = ns["bench"](lib, sh)
        row.append(f"{ms*1e3:9.1f}")
    print(str(sh).ljust(44) + "".join(row), flush=True)
PY
cat gpurun_out/exp6_debug.log
