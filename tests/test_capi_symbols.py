"""The C-ABI library loads on a box without a GPU and exports every symbol that
include/coxgraph_b200.h declares; compute entry points fail loudly without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "coxgraph_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from coxgraph_b200 import capi
    lib = capi.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in the header but not exported"
        assert n in capi.SYMBOLS, f"{n} has no ctypes prototype in capi.SYMBOLS"
    assert set(capi.SYMBOLS) <= set(names), "capi.SYMBOLS binds symbols the header does not declare"


def test_struct_layouts_match_the_header():
    from coxgraph_b200 import capi
    from oracle import oracle_py as orc
    assert C.sizeof(capi.IntegratorConfig) == 15 * 4
    assert [f[0] for f in capi.IntegratorConfig._fields_] == [f[0] for f in orc.IntegratorConfig._fields_]
    assert C.sizeof(capi.IntegrateStats) == 56 and C.sizeof(capi.MergeStats) == 24
    assert C.sizeof(capi.ReprojectStats) == 48 and C.sizeof(capi.HashStats) == 48
    assert C.sizeof(capi.StageProfile) == 48
    # the structs as the C compiler lays them out (the header is plain C)
    import subprocess
    import sys
    import tempfile
    src = ('#include <stdio.h>\n#include "coxgraph_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
           'sizeof(cg_integrator_config),sizeof(cg_integrate_stats),sizeof(cg_merge_stats),'
           'sizeof(cg_reproject_stats),sizeof(cg_hash_stats),sizeof(cg_stage_profile));return 0;}')
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "s.c"), "w") as f:
            f.write(src)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"),
                               os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(t) for t in (capi.IntegratorConfig, capi.IntegrateStats,
                                           capi.MergeStats, capi.ReprojectStats, capi.HashStats,
                                           capi.StageProfile)], sizes
    del sys
    assert capi.PACKED_BLOCK_BYTES == 16 + 4096 * 12


def test_defaults_are_upstream_voxblox_defaults():
    from coxgraph_b200 import TsdfIntegratorConfig
    c = TsdfIntegratorConfig()
    assert abs(c.default_truncation_distance - 0.1) < 1e-7 and c.max_weight == 10000.0
    assert (c.voxel_carving_enabled, c.allow_clear, c.use_weight_dropoff) == (1, 1, 1)
    assert (c.use_const_weight, c.enable_anti_grazing, c.method) == (0, 0, 1)
    assert abs(c.min_ray_length_m - 0.1) < 1e-7 and c.max_ray_length_m == 5.0
    with pytest.raises(AttributeError):
        TsdfIntegratorConfig(no_such_field=1)


def test_host_side_helpers_need_no_gpu():
    from coxgraph_b200 import capi
    lib = capi.load()
    assert lib.cg_version().decode().startswith("coxgraph_b200")
    owners = {lib.cg_block_owner(x, y, z, 8) for x in range(-4, 4) for y in range(-4, 4)
              for z in range(-2, 2)}
    assert owners == set(range(8))
    assert lib.cg_block_owner(1, 2, 3, 1) == 0 and lib.cg_block_owner(1, 2, 3, 0) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from coxgraph_b200 import Context, capi
    with pytest.raises(capi.CgError) as e:
        Context(0)
    assert e.value.status == capi.CG_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "coxgraph_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "tsdf_oracle", "oracle_py",
                               "oracle/", "liboracle", "orc_"):
                    assert needle not in src, f"{f} references the oracle ({needle})"
