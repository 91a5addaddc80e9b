"""The C-ABI library loads on a machine without a GPU and exports every symbol include/vqa_b200.h declares;
op field tables of both sides agree; errors come back as codes + messages (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from conftest import REPO
from vqa_b200 import program as P
from vqa_b200 import runtime as R
from vqa_b200.build import CSRC, build


@pytest.fixture(scope="module")
def lib():
    build()
    return R.lib()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(REPO, "include", "vqa_b200.h")).read()
    names = set(re.findall(r"\b(vqa_[a-z_0-9]+)\s*\(", header))
    assert {"vqa_plan_create", "vqa_plan_run", "vqa_plan_run_range", "vqa_plan_destroy", "vqa_last_error",
            "vqa_abi_version", "vqa_device_check", "vqa_op_num_fields", "vqa_launch_count",
            "vqa_plan_num_launches", "vqa_plan_op_kernel_name", "vqa_resize_bilinear_u8",
            "vqa_accuracy_update"} <= names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.vqa_abi_version() == R.ABI_VERSION


def test_field_tables_agree_and_header_is_generated(lib):
    for kind, code in P.KINDS.items():
        ni, np_, nf = C.c_int(), C.c_int(), C.c_int()
        assert lib.vqa_op_num_fields(code, C.byref(ni), C.byref(np_), C.byref(nf)) == 0
        assert (ni.value, np_.value, nf.value) == tuple(len(P.FIELDS[kind][k]) for k in "ipf")
        assert len(P.FIELDS[kind]["i"]) <= R.OP_NI and len(P.FIELDS[kind]["p"]) <= R.OP_NP
    assert open(os.path.join(CSRC, "op_fields.h")).read() == P.generate_fields_header() + "\n"
    assert C.sizeof(R.VqaOp) == 8 + 4 * R.OP_NI + 4 * R.OP_NF + 8 * R.OP_NP


def test_errors_are_codes_not_aborts(lib):
    assert lib.vqa_op_num_fields(999, None, None, None) == -1
    assert b"unknown op kind" in lib.vqa_last_error()
    h = C.c_void_p()
    assert lib.vqa_plan_create(None, 0, 0, C.byref(h)) == -1 and not h.value
    assert lib.vqa_plan_run(None, None, 0, None) == -1
    lib.vqa_plan_destroy(None)   # no-op


def test_no_cpu_fallback():
    import torch
    from vqa_b200 import VQAModel
    m = VQAModel(vocab_size=50, num_answers=7).eval()
    with pytest.raises(Exception) as e:
        m(torch.zeros(1, 3, 224, 224), torch.zeros(1, 20, dtype=torch.long))
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)
    with pytest.raises(NotImplementedError):
        m.train()(torch.zeros(1, 3, 224, 224), torch.zeros(1, 20, dtype=torch.long))
