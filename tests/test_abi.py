"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the Python operator mirrors the reference constructor contract.  No compute calls."""
import os
import re

import pytest
import torch

import egc_b200
from egc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "egc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(egc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = egc_b200.load()
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/egc_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in egc_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_build_info():
    lib = egc_b200.load()
    assert lib.egc_abi_version() == _lib.ABI_VERSION == 5
    info = lib.egc_build_info().decode()
    assert "sm_100a" in info and "chunk_edges=256" in info


def test_struct_layouts_match_header():
    assert _lib.LayerDesc.aggr.offset == 24 and _lib.LayerDesc.sigmoid.offset == 24 + 4 * 8
    import ctypes
    assert ctypes.sizeof(_lib.LayerDesc) == 64 and _lib.LayerDesc.relu.offset == 60
    assert ctypes.sizeof(_lib.RowPlan) == 8 + 4 * 8


def test_constructor_contract_matches_reference():
    # ref optimized_layers.py:89-94 error behaviour, :105-113 parameter shapes, :280-286 repr
    with pytest.raises(ValueError, match="divisible"):
        egc_b200.EGConv(16, 100, num_heads=8)
    with pytest.raises(ValueError, match="Unsupported aggregator"):
        egc_b200.EGConv(16, 32, aggrs=["symadd"])
    conv = egc_b200.EGConv(128, 128, aggrs=["symnorm", "max", "std"], num_heads=4, num_bases=4)
    sd = conv.state_dict()
    assert list(sd) == ["bases_weight", "bias", "comb_weight.weight", "comb_weight.bias"]
    assert sd["bases_weight"].shape == (128, 128) and sd["comb_weight.weight"].shape == (48, 128)
    assert sd["comb_weight.bias"].shape == (48,) and sd["bias"].shape == (128,)
    assert repr(conv) == "EGConv(128, 128, ['symnorm', 'max', 'std'])"
    assert float(sd["bias"].abs().max()) == 0.0
    bound = (6.0 / (128 + 128)) ** 0.5
    assert float(sd["bases_weight"].abs().max()) <= bound
    assert egc_b200.EGConv(8, 16, bias=False).bias is None
    # known-answer parameter count: one ZINC EGC-S layer (hidden 168, H8 B4) = 19 688
    # (4 layers = 78 752 of the 102 861 printed at output/pretrained.txt:41)
    zinc = egc_b200.EGConv(168, 168, aggrs=["symnorm"], num_heads=8, num_bases=4)
    assert sum(p.numel() for p in zinc.parameters()) == 19688


def test_no_cpu_fallback():
    conv = egc_b200.EGConv(8, 16, num_heads=4)
    with pytest.raises(RuntimeError, match="CUDA"):
        conv(torch.randn(4, 8), torch.zeros(2, 3, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        egc_b200.GraphStructure.from_edge_index(torch.zeros(2, 3, dtype=torch.long), 4, True, True)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "egc_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src, f"egc_b200/{fn} mentions the oracle"


def test_adjacency_inputs_accepted_without_torch_sparse():
    """the layer accepts the repo's SparseTensor, a torch sparse CSR tensor, and any object with
    .csr() / .sparse_sizes() (a real torch_sparse.SparseTensor); everything else is a TypeError."""
    import warnings

    from egc_b200.graph import adjacency_to_csr
    rowptr, col, val = torch.tensor([0, 2, 3]), torch.tensor([0, 1, 1]), torch.tensor([1., 2., 3.])
    own = egc_b200.SparseTensor(rowptr=rowptr, col=col, value=val, sparse_sizes=(2, 3))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        native = torch.sparse_csr_tensor(rowptr, col, val, size=(2, 3))

    class Foreign:
        def csr(self):
            return rowptr, col, None

        def sparse_sizes(self):
            return (2, 3)

    for adj, has_val in ((own, True), (native, True), (Foreign(), False)):
        rp, c, v, n_src = adjacency_to_csr(adj)
        assert torch.equal(rp, rowptr) and torch.equal(c, col) and n_src == 3 and (v is not None) == has_val
    with pytest.raises(TypeError):
        adjacency_to_csr(object())
    # unsorted COO input is sorted by (row, col) with values carried along; .t() swaps the roles and stays sorted
    coo = egc_b200.SparseTensor(row=torch.tensor([1, 0, 0]), col=torch.tensor([1, 1, 0]), value=torch.tensor([3., 2., 1.]),
                                sparse_sizes=(2, 2))
    assert coo.coo()[0].tolist() == [0, 0, 1] and coo.coo()[1].tolist() == [0, 1, 1] and coo.coo()[2].tolist() == [1., 2., 3.]
    t = coo.t()
    assert t.coo()[0].tolist() == [0, 1, 1] and t.coo()[1].tolist() == [0, 0, 1] and t.coo()[2].tolist() == [1., 2., 3.]
    sym = egc_b200.SparseTensor(row=torch.tensor([0]), col=torch.tensor([1]), sparse_sizes=(2, 2)).to_symmetric()
    assert sym.coo()[0].tolist() == [0, 1] and sym.coo()[1].tolist() == [1, 0]


def test_plain_c_program_links_against_the_abi(tmp_path):
    """include/egc_b200.h is a C header: a C99 translation unit includes it, links libegc_b200.so and calls the
    device-free entry points (no GPU needed)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi_probe.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "egc_b200.h"
int main(void) {
  egc_layer_desc d;
  memset(&d, 0, sizeof d);
  d.n_dst = d.n_src = 10; d.heads = 4; d.bases = 4; d.dim = 32; d.n_aggr = 3;
  d.aggr[0] = EGC_AGGR_SYMNORM; d.aggr[1] = EGC_AGGR_MAX; d.aggr[2] = EGC_AGGR_STD;
  printf("%d %d %d %s\n", egc_abi_version(), (int)egc_saved_slots(&d), (int)egc_saved_arg_slots(&d), egc_build_info());
  /* a call that must fail cleanly without a device: null descriptor */
  int rc = egc_aggregate_fwd(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL);
  printf("%d %s\n", rc, egc_last_error_string());
  return 0;
}
''')
    exe = tmp_path / "abi_probe"
    libdir = os.path.join(ROOT, "egc_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-legc_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    version, saved, saved_arg, *info = out[0].split()
    assert int(version) == _lib.ABI_VERSION and int(saved) == 4 and int(saved_arg) == 1 and "sm_100a" in " ".join(info)
    rc, msg = out[1].split(" ", 1)
    assert int(rc) == -1 and "null descriptor" in msg
