"""CPU checks of the C-ABI boundary: the shared library loads, exports every function include/calm_b200.h declares, the
ctypes prototype table names exactly those functions, and the ctypes struct mirrors have the C compiler's layout.
No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess

import pytest

import calm_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "calm_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return sorted(set(re.findall(r"\b(calm_[a-z0-9_]+)\s*\(", src)))


def test_header_and_prototype_table_agree():
    assert _declared_functions() == sorted(calm_lib.PROTOTYPES)


def test_library_loads_and_exports_every_declared_symbol():
    lib = calm_lib.load()                      # builds with nvcc when the .so is absent; raises if it cannot
    assert lib.calm_abi_version() == 1
    raw = C.CDLL(calm_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(raw, name), name


def test_ctypes_structs_match_the_c_layout(tmp_path):
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "calm_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(calm_gemm_args), sizeof(calm_sn_layer), '
                    'sizeof(calm_sn_item), sizeof(calm_trainer_step_args), offsetof(calm_gemm_args, alpha), '
                    'offsetof(calm_sn_layer, item_count), offsetof(calm_trainer_step_args, use_scaler));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(calm_lib.GemmArgs), C.sizeof(calm_lib.SnLayer), C.sizeof(calm_lib.SnItem), C.sizeof(calm_lib.TrainerStepArgs),
            calm_lib.GemmArgs.alpha.offset, calm_lib.SnLayer.item_count.offset, calm_lib.TrainerStepArgs.use_scaler.offset]
    assert got == want


def test_trainer_glue_has_no_cpu_fallback():
    import torch
    import calm_trainer
    p = torch.nn.Parameter(torch.zeros(8))
    with pytest.raises(calm_lib.CalmError):
        calm_trainer.TrainerStep([p])
    with pytest.raises(calm_lib.CalmError):
        calm_trainer.soft_target_cross_entropy(torch.zeros(2, 5, requires_grad=True), torch.full((2, 5), 0.2))
