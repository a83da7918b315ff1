"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/fastkv_b200.h declares; host-side control flow that needs no kernel."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fastkv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fkv_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from fastkv_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fastkv_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(names)
    assert lib.fkv_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_c_layout():
    """ctypes mirrors must have the C sizes (checked against sizes compiled from the header)."""
    import subprocess, tempfile
    from fastkv_b200 import _lib
    src = '#include <stdio.h>\n#include "fastkv_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(fkv_view4),sizeof(fkv_score_params),sizeof(fkv_topk_params),sizeof(fkv_gather_kv_params),' \
          'sizeof(fkv_gather_rows_params),sizeof(fkv_decode_attn_params),sizeof(fkv_cache_append_params),' \
          'sizeof(fkv_compress_params),sizeof(fkv_select_compact_params),sizeof(fkv_exchange_params),sizeof(fkv_allreduce_params),(size_t)FKV_WORKSPACE_HEADER_BYTES);return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mirrors = [_lib.View4, _lib.ScoreParams, _lib.TopkParams, _lib.GatherKvParams, _lib.GatherRowsParams,
               _lib.DecodeAttnParams, _lib.CacheAppendParams, _lib.CompressParams, _lib.SelectCompactParams, _lib.ExchangeParams, _lib.AllreduceParams]
    assert sizes[:-1] == [ctypes.sizeof(m) for m in mirrors]
    assert sizes[-1] == _lib.WS_HEADER                         # the persistent region the Python side reserves


def test_strerror_and_error_mapping():
    from fastkv_b200 import _lib
    lib = _lib.load()
    assert b"Pooling method not supported" in lib.fkv_strerror(_lib.EPOOL)
    with pytest.raises(ValueError, match="Pooling method not supported"):
        _lib.check(_lib.EPOOL)
    with pytest.raises(ValueError, match="Pooling method not supported"):
        _lib.pool_code("minpool")
    with pytest.raises(RuntimeError):
        _lib.check(2)          # cudaErrorMemoryAllocation


def test_null_and_bad_arguments_return_codes_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    from fastkv_b200 import _lib
    lib = _lib.load()
    assert lib.fkv_score(None, None) == -1
    p = _lib.TopkParams()
    p.struct_size = ctypes.sizeof(_lib.TopkParams)
    p.rows, p.n, p.k = 1, 8, 9
    p.src, p.out = 16, 16
    assert lib.fkv_topk(ctypes.byref(p), None) == -5           # k > n
    p.struct_size = 4
    assert lib.fkv_topk(ctypes.byref(p), None) == -1
    sp = _lib.ScoreParams()
    sp.struct_size = ctypes.sizeof(_lib.ScoreParams)
    sp.B, sp.Hq, sp.Hkv, sp.L, sp.D = 1, 8, 2, 64, 96
    assert lib.fkv_score(ctypes.byref(sp), None) == -3          # head_dim 96
    sp.D, sp.kernel_size = 128, 4
    assert lib.fkv_score(ctypes.byref(sp), None) == -4          # even kernel_size
    assert lib.fkv_workspace_bytes(0, 1, 32, 8, 32768, 8, 2048, 2040) > _lib.WS_HEADER
    assert lib.fkv_topk_workspace_bytes(1, 1000, 10) > _lib.WS_HEADER and lib.fkv_decode_workspace_bytes(1, 32, 128, 2304) > _lib.WS_HEADER
    sc = _lib.SelectCompactParams()
    assert lib.fkv_select_compact(ctypes.byref(sc), None) == -1     # struct_size guard
    assert lib.fkv_select_compact(None, None) == -1
    assert lib.fkv_workspace_status(None, None) == -7               # FKV_EWORKSPACE
    xp = _lib.ExchangeParams()
    assert lib.fkv_exchange_push(ctypes.byref(xp), None) == -1      # struct_size guard
    xp.struct_size = ctypes.sizeof(_lib.ExchangeParams)
    xp.rank, xp.world, xp.B, xp.heads_local, xp.P, xp.score = 0, 9, 1, 4, 1024, 16
    assert lib.fkv_exchange_push(ctypes.byref(xp), None) == -1      # world > FKV_MAX_PEERS
    xp.world = 2
    xp.region[0] = 4096
    assert lib.fkv_exchange_wait(ctypes.byref(xp), None) == -1      # region of rank 1 missing
    xp.region[1] = 8192 + 8
    assert lib.fkv_exchange_push(ctypes.byref(xp), None) == -6      # unaligned region
    xp.region[1] = 8192
    xp.region_bytes = _lib.EXCHANGE_HEADER + 2 * 4 * 1024 * 4 - 1
    assert lib.fkv_exchange_push(ctypes.byref(xp), None) == -7      # region too small for the gathered rows
    assert lib.fkv_exchange_status(None, None) == -1
    ar = _lib.AllreduceParams()
    assert lib.fkv_allreduce_sum(ctypes.byref(ar), None) == -1       # struct_size guard
    ar.struct_size = ctypes.sizeof(_lib.AllreduceParams)
    ar.rank, ar.world, ar.dtype, ar.n, ar.data = 0, 2, 0, 4096, 4096
    ar.region[0], ar.region[1] = 8192, 16384
    ar.region_bytes = _lib.ALLREDUCE_HEADER + 2 * 2 * 8192
    assert lib.fkv_allreduce_sum(ctypes.byref(ar), None) == -7       # 8 KiB of bf16 do not fit half of an 8 KiB slot
    ar.n = 4100
    assert lib.fkv_allreduce_sum(ctypes.byref(ar), None) == -6       # byte count not a multiple of 16
    ar.n, ar.dtype = 4096, 7
    assert lib.fkv_allreduce_sum(ctypes.byref(ar), None) == -2
    assert b"peer" in lib.fkv_strerror(_lib.ETIMEOUT)
    sp.D, sp.kernel_size, sp.hist_slot = 128, 7, 3
    sp.pooling, sp.window_size, sp.h1 = 0, 8, 2
    assert lib.fkv_score(ctypes.byref(sp), None) in (-1, -6, -7)    # rejected before any launch (null views / workspace / bad slot)


def test_early_return_is_by_reference_and_needs_no_gpu():
    """baselines/fastkv/utils.py:89-91: q_len < max_capacity_prompt returns the inputs themselves."""
    from fastkv_b200 import FastKVCluster, SnapKVCluster
    k, v = torch.randn(1, 2, 64, 128), torch.randn(1, 2, 64, 128)
    q = torch.randn(1, 8, 64, 128)
    cl = FastKVCluster(max_capacity_prompt=128, tsp_layer=True, tsp_length=32)
    ko, vo, tsp = cl.update_kv(k, q, v, None, 4, 0)
    assert ko is k and vo is v and tsp is None
    sn = SnapKVCluster(window_size=8, max_capacity_prompt=128)
    ko, vo = sn.update_kv(k, q, v, None, 4)
    assert ko is k and vo is v


def test_proportional_mode_rewrites_attributes_like_the_reference():
    """utils.py:86-87: proportional mode overwrites self.max_capacity_prompt with int(q_len * retain_rate)."""
    from fastkv_b200 import FastKVCluster
    cl = FastKVCluster(window_size=8, max_capacity_prompt=512, retain_rate=2.0, eviction_mode="proportional")
    k, v, q = torch.randn(1, 2, 50, 128), torch.randn(1, 2, 50, 128), torch.randn(1, 8, 50, 128)
    ko, vo, tsp = cl.update_kv(k, q, v, None, 4, 0)          # C = 100 > q_len -> early return
    assert cl.max_capacity_prompt == 100 and ko is k and tsp is None


def test_no_cpu_fallback():
    from fastkv_b200 import FastKVCluster
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    cl = FastKVCluster(window_size=8, max_capacity_prompt=16)
    k, v, q = torch.randn(1, 2, 64, 128), torch.randn(1, 2, 64, 128), torch.randn(1, 8, 64, 128)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cl.update_kv(k, q, v, None, 4, 0)


def test_constructor_assert_and_reset_semantics():
    from fastkv_b200 import FastKVCluster
    with pytest.raises(AssertionError):
        FastKVCluster(window_size=8, max_capacity_prompt=8)      # utils.py:54
    cl = FastKVCluster()
    cl.reset(tsp_length=1024)
    assert cl.tsp_length == 1024          # not the reference's accidental 1-tuple (utils.py:73)


_C_CLIENT = r"""
#include <stdio.h>
#include <string.h>
#include "fastkv_b200.h"
int main(void) {
  fkv_score_params sp; memset(&sp, 0, sizeof sp);
  sp.struct_size = (uint32_t)sizeof sp;
  printf("abi %d\n", fkv_abi_version());
  printf("ws %zu\n", fkv_workspace_bytes(FKV_MODE_FASTKV, 1, 32, 8, 32768, 8, 2048, 2040));
  printf("rc %d\n", fkv_score(&sp, NULL));   /* invalid arguments: rejected before anything touches the GPU */
  return 0;
}
"""


def test_header_is_plain_c_and_links_from_a_c_program(tmp_path):
    """include/fastkv_b200.h must be usable from the reference-side FFI of any language: compile a C99 client with
    -pedantic against it, link the shared library, call the query entry points (no GPU involved)."""
    import os, shutil, subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from fastkv_b200 import _lib
    _lib.load()                                                   # builds the library if needed
    libdir = os.path.join(root, "fastkv_b200")
    src = tmp_path / "client.c"
    src.write_text(_C_CLIENT)
    exe = tmp_path / "client"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src),
                        "-o", str(exe), "-L", libdir, "-l:libfastkv_b200.so", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    lines = dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())
    assert int(lines["abi"]) == _lib.load().fkv_abi_version() and int(lines["ws"]) > 0 and int(lines["rc"]) < 0
