"""CPU tests: the in-tree LA / BLA table builders (fractalshark_b200/csrc/host/fs_host.cpp, the inputs of every LAv2 /
BLA parity case and of the headline bench) against the REFERENCE's own builders -- LAReference.cpp + LAInfoDeep.h and
BLAS.cpp compiled from /root/reference into oracle/_ref/libref_host.so (oracle/Makefile, ref_host_harness.cpp) -- on
the same orbit.  Byte for byte: LAInfoDeep[], LAStageInfo[], ATInfo, BLA<T> levels."""
import ctypes as C

import numpy as np
import pytest

import cases
import ref_host
from fractalshark_b200 import RenderAlgorithm as A

pytestmark = pytest.mark.skipif(not ref_host.available(),
                                reason="oracle/_ref/libref_host.so (the reference's host builders) not built")


def _at_bytes(la):
    d = la.descriptor()
    return np.frombuffer((C.c_ubyte * la.at_bytes).from_address(d.at), dtype=np.uint8).copy()


@pytest.mark.parametrize("threading", [1, 0], ids=["single_threaded", "reference_default_threading"])
@pytest.mark.parametrize("view_id,iter_bytes", [(1, 4), (1, 8), (5, 4), (5, 8), (19, 4), (14, 4), (14, 8), (100, 4)])
def test_la_table_is_the_one_the_reference_builds(built, view_id, iter_bytes, threading):
    """LAReference<IterType, HDRFloat<float>, float, Disable>::GenerateApproximationData (LAReference.cpp:971-1017: stage 0
    :28-207 or its multi-threaded form :215-771, higher stages :774-968, AT :1050-1074) on the generator's orbit.  View 14:
    33,844 records in 5 stages with an AT block -- the table behind the headline's skip factor."""
    _, coords, orbit, la, n = cases.make_inputs(view_id, 192, 108, A.GpuHDRx32PerturbedLAv2, None, iter_bytes)
    ref = ref_host.RefLaTable(orbit, iter_bytes, n, threading)
    assert (ref.num_las, ref.stage_count, ref.use_at, ref.is_valid) == (la.num_las, la.stage_count, la.use_at, la.is_valid)
    assert ref.las_elem_bytes == la.las_elem_bytes and ref.at_bytes == la.at_bytes
    mine, theirs = la.las_numpy().copy(), ref.las.copy()
    if iter_bytes == 8:
        # LAInfoDeep<uint64_t, ...>: 4 indeterminate padding bytes between MinMag (ends at 60) and LAi (at 64)
        mine[:, 60:64] = 0
        theirs[:, 60:64] = 0
    np.testing.assert_array_equal(mine, theirs)
    # the reference keeps MaxLAStages (1024) stage slots; the first LAStageCount are the table
    np.testing.assert_array_equal(la.stages_numpy()[: la.stage_count], ref.stages[: ref.stage_count])
    np.testing.assert_array_equal(_at_bytes(la), ref.at)
    if view_id == 14:
        assert la.num_las == 33844 and la.stage_count == 5 and la.use_at


@pytest.mark.parametrize("view_id", [1, 5, 14, 100])
def test_bla_table_is_the_one_the_reference_builds(built, view_id):
    """BLAS<IterType, HDRFloat<float>>::Init(count, MaxRadius) (BLAS.cpp:212-254; leaves :74-92, merge :25-47) on the
    generator's orbit: same level count, LM2 and every BLA<T> record."""
    _, coords, orbit, bl, n = cases.make_inputs(view_id, 192, 108, A.GpuHDRx32PerturbedBLA, None, 4)
    ref = ref_host.RefLaTable(orbit, 4, n, 1, with_blas=True)
    assert len(ref.blas_levels) == bl.num_levels and ref.blas_lm2 == bl.lm2
    for lv in range(bl.num_levels):
        np.testing.assert_array_equal(bl.level_numpy(lv), ref.blas_levels[lv])
