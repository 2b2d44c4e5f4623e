"""The compile-time bound bookkeeping of the inverse (GS) stages in aloha_b200/csrc/ntt_kernels.cu
(ALOHA_GS_REDUCE_PAIR / ALOHA_GS_STAGE / ALOHA_GS_NORMALISE / ALOHA_GS_LAST) restated in Python: for both
arithmetic forms and every block shape the kernels instantiate, no intermediate can reach 16q (= the 64-bit
word for q < 2^60) and the subtraction offset always covers the subtrahend -- the condition behind the
kernels' `__trap()` that must fold away."""
import pytest

FORMS = {"generic": dict(MULB=2, REDB=8), "pm": dict(MULB=3, REDB=2)}


def off_units(b):
    return 2 if b <= 2 else 4 if b <= 4 else 8


def reduce_pair(bnd, i, j, REDB):
    if bnd[i] + bnd[j] > 16:
        if bnd[i] > REDB:
            bnd[i] = REDB
        if bnd[j] > REDB:
            bnd[j] = REDB
    assert bnd[i] + bnd[j] <= 16, "sum would pass 16q"
    by = bnd[j]
    assert off_units(by) >= by, "offset smaller than the subtrahend"
    assert bnd[i] + off_units(by) <= 16, "x - y + off would pass 16q"
    return by


def gs_stage(bnd, nelem, half, MULB, REDB):
    for e in range(nelem):
        if e & half:
            continue
        by = reduce_pair(bnd, e, e + half, REDB)
        bnd[e] += by
        bnd[e + half] = MULB


def normalise(bnd):
    for e in range(len(bnd)):
        assert bnd[e] <= 16
        bnd[e] = 2


@pytest.mark.parametrize("form", FORMS)
def test_inverse_row_pass_bounds(form):
    F = FORMS[form]
    for last_in_rows in (False, True):          # S1 == 0: the row pass holds the transform's last stage
        bnd = [2] * 16
        for lt in range(4):
            gs_stage(bnd, 16, 1 << lt, **F)
        normalise(bnd)
        for lt in range(4, 8):
            if last_in_rows and lt == 7:
                for i in range(8):
                    reduce_pair(bnd, i, i + 8, F["REDB"])
            else:
                gs_stage(bnd, 16, 1 << (lt - 4), **F)
        assert max(bnd) <= 16


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("s1", range(1, 9))
def test_inverse_column_pass_bounds(form, s1):
    F = FORMS[form]
    la = min(s1, 4)
    lb = s1 - la
    e = 1 << la
    bnd = [2] * e
    if lb:
        for b in range(lb):
            gs_stage(bnd, e, 1 << b, **F)
        normalise(bnd)
    for b in range(lb, s1):
        if b == s1 - 1:
            for i in range(e // 2):
                reduce_pair(bnd, i, i + e // 2, F["REDB"])
        else:
            gs_stage(bnd, e, 1 << (b - lb), **F)


@pytest.mark.parametrize("form,grow,redb", [("generic", 2, 8), ("pm", 3, 2)])
def test_forward_bounds(form, grow, redb):
    """cols_out_bound + the row pass: one scalar bound, an upper-input reduction when the next stage would pass 16."""
    for s1 in range(0, 9):
        b = 2
        for _ in range(s1 + 8):
            if b + grow > 16:
                b = redb
            b += grow
            assert b <= 16
