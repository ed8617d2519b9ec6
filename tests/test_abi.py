"""CPU: the C-ABI library loads and exports exactly what include/tt_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "tt_b200.h"


def declared_functions():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for n in ["tt_pool_weighted", "tt_pool_weighted_gather", "tt_attention_logits", "tt_pool_attention",
              "tt_pool_attention_gather", "tt_flat_build", "tt_flat_search", "tt_flat_search_exact", "tt_topk_merge",
              "tt_flat_search_workspace_bytes", "tt_last_error", "tt_abi_version"]:
        assert n in names


def test_library_exports_every_declared_symbol(native_lib):
    for name in declared_functions():
        assert hasattr(native_lib, name), f"{name} declared in include/tt_b200.h but not exported"


def test_binding_covers_header(native_lib):
    from two_tower_model_v2_b200 import _native
    assert sorted(_native.SIGNATURES) == declared_functions()
    assert native_lib.tt_abi_version() == _native.ABI_VERSION


def test_argument_errors_are_reported_without_a_gpu(native_lib):
    from two_tower_model_v2_b200 import _native
    rc = native_lib.tt_pool_weighted(None, None, None, 1, 1, 1, None)
    assert rc == 1 and "null pointer" in _native.last_error()
    with pytest.raises(ValueError):
        _native.check(rc, "tt_pool_weighted")
    assert native_lib.tt_flat_pitch(384) == 384 and native_lib.tt_flat_pitch(100) == 128
    assert native_lib.tt_flat_search_workspace_bytes(10_000_000, 384, 128, 100) > 0
    assert native_lib.tt_flat_search_workspace_bytes(0, 384, 1, 1) == 0


def test_sass_is_blackwell_native():
    """The scan kernel must contain tcgen05 MMA / TMEM loads / TMA loads (SASS: UTCHMMA, LDTM, UTMALDG)."""
    import shutil
    import subprocess
    from two_tower_model_v2_b200 import _native
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(_native.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
    assert "HMMA.16" not in sass, "legacy mma.sync path found"


def test_torch_ops_are_cuda_only():
    """Every op is registered under torch.ops.tt and has no CPU (or Meta) implementation: CPU tensors fail loudly."""
    import pytest
    import torch
    import two_tower_model_v2_b200  # noqa: F401
    for name in ("pool_weighted", "pool_weighted_gather", "attention_logits", "pool_attention", "pool_attention_gather",
                 "topk_merge", "flat_build", "flat_search", "flat_search_exact"):
        assert hasattr(torch.ops.tt, name), name
    with pytest.raises(NotImplementedError):
        torch.ops.tt.pool_weighted(torch.zeros(1, 2, 4), torch.ones(1, 2))
    with pytest.raises(NotImplementedError):
        torch.ops.tt.flat_search(torch.zeros(2, 8), torch.zeros(4, 8), torch.zeros(4, 64, dtype=torch.bfloat16),
                                 torch.zeros(4), 2, 0)


def test_header_is_plain_c_and_links_from_c(tmp_path, native_lib):
    """include/tt_b200.h must be usable from C (the boundary is a C-ABI): compile a C99 translation unit against it
    with -Wall -Werror -pedantic, link it to the shared library and run it without a GPU (argument errors and
    host-only helpers must work; nothing may crash)."""
    import shutil
    import subprocess
    from two_tower_model_v2_b200 import _native
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi_demo.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "tt_b200.h"
int main(void) {
  int32_t plan[16];
  if (tt_abi_version() != TT_B200_ABI_VERSION) return 10;
  if (tt_flat_pitch(384) != 384 || tt_flat_pitch(1) != 64) return 11;
  if (tt_pool_weighted(NULL, NULL, NULL, 1, 1, 1, NULL) != TT_ERR_INVALID) return 12;
  if (strstr(tt_last_error(), "null pointer") == NULL) return 13;
  if (tt_flat_plan_describe(10000000, 384, 4096, 100, plan) != TT_OK) return 14;
  if (plan[0] != 1 || plan[1] != 256 || plan[2] != 6) return 15;         /* supported, CTA pairs, 6 K-blocks */
  if (tt_flat_search_workspace_bytes(10000000, 384, 4096, 100) == 0) return 16;
  if (tt_flat_shard_plan_ok(1250000, 10000000, 384, 4096, 100) != 1) return 17;
  if (tt_shard_merge(NULL, 0, 0, 0, 0, 0, 1, 1, 1, NULL, NULL, NULL, NULL, NULL) != TT_ERR_INVALID) return 18;
  printf("abi ok, TT_FLAT_MAX_K=%d TT_SHARD_TOPR=%d\n", TT_FLAT_MAX_K, TT_SHARD_TOPR);
  return 0;
}
''')
    exe = tmp_path / "abi_demo"
    lib = Path(_native.LIB_PATH)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", str(HEADER.parent), str(src), "-o", str(exe),
                        "-L", str(lib.parent), "-ltt_b200", f"-Wl,-rpath,{lib.parent}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, f"exit {run.returncode}: {run.stdout} {run.stderr}"
    assert "abi ok" in run.stdout
