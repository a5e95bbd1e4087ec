import sys, time, os, ctypes as C
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common
pkg = common.pkg
be = pkg.get_backend(); be.set_table_path(common.table_dir())
os.environ["B200_SKIP_SCRATCH_OUTPUTS"] = "1"
hii, dim = int(sys.argv[1]), int(sys.argv[2])
inputs = common.make_inputs(hii=hii, dim=dim, box_len=1.5 * hii)
be.lib.b200_profile_report.argtypes = [C.c_char_p, C.c_int]
be.lib.b200_profile_enable(1)
t0 = time.perf_counter()
ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
t1 = time.perf_counter()
buf = C.create_string_buffer(1 << 16)
be.lib.b200_profile_report(buf, len(buf))
print(f"ICs HII={hii} DIM={dim}: wall {t1-t0:.2f} s")
rows = []
for ln in buf.value.decode().splitlines():
    nm, cnt, tot = ln.rsplit(None, 2)
    rows.append((float(tot), int(cnt), nm))
for tot, cnt, nm in sorted(rows, reverse=True)[:12]:
    print(f"  {nm:40s} {cnt:5d} launches {tot:10.1f} ms")
