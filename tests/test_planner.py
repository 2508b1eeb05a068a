"""The planner's decisions (jpgpu_plan_info: host only, no GPU): subsequence length, look-back, restart-interval mode.
The numbers are the ones DESIGN.md section 4.2 gives reasons for; a change of heuristics has to change this file too."""
import pytest

from jpeg_rust_b200 import EXT_DRI, plan_info, synth


@pytest.fixture(scope="module")
def one_1080p_420():
    return synth.synth_jpeg(1, 1920, 1080, "420")


def test_batch_that_fills_the_machine(one_1080p_420):
    p = plan_info([one_1080p_420], copies=1024)        # BASELINE configs[2]
    assert (p["sub_bits"], p["lookback_bits"], p["seg_bits"], p["write_parts"], p["groups"]) == (8192, 1024, 1024, 1, 3)
    assert p["interval_images"] == 0 and p["warp_jobs"] % 8 == 0


def test_mid_size_batches_trade_threads_for_less_overhead(one_1080p_420):
    p512, p256 = plan_info([one_1080p_420], copies=512), plan_info([one_1080p_420], copies=256)
    assert (p512["sub_bits"], p512["lookback_bits"]) == (4096, 2048)
    assert (p256["sub_bits"], p256["lookback_bits"]) == (4096, 2048)
    assert p256["groups"] == 3 and plan_info([one_1080p_420], copies=150)["groups"] == 2


def test_small_420_batches_look_back_further(one_1080p_420):
    assert plan_info([one_1080p_420], copies=64)["lookback_bits"] == 2048     # round 2: the repair walks got cheaper (multi-symbol tables)
    p1 = plan_info([one_1080p_420])
    assert (p1["sub_bits"], p1["lookback_bits"], p1["groups"]) == (1024, 8192, 1)


@pytest.mark.parametrize("sub", ["444", "422", "gray", "440"])
def test_mcus_of_up_to_four_blocks_keep_the_short_look_back(sub):
    assert plan_info([synth.synth_jpeg(2, 640, 480, sub)])["lookback_bits"] == 1024


def test_restart_interval_mode_is_chosen_per_image():
    dense = synth.synth_jpeg(3, 1024, 768, "444", restart_interval=4)       # ~600 bits per interval
    sparse = synth.synth_jpeg(3, 1024, 768, "444", restart_interval=128)    # one MCU row: far longer than a subsequence
    plain = synth.synth_jpeg(3, 1024, 768, "444")
    assert plan_info([dense, sparse, plain], ext=EXT_DRI)["interval_images"] == 1
    assert plan_info([dense], ext=EXT_DRI, copies=5)["interval_images"] == 5


def test_environment_overrides(monkeypatch, one_1080p_420):
    monkeypatch.setenv("JPGPU_SUBSEQ_BITS", "16384")
    monkeypatch.setenv("JPGPU_LOOKBACK_BITS", "300")
    monkeypatch.setenv("JPGPU_WRITE_PARTS", "4")
    monkeypatch.setenv("JPGPU_GROUPS", "2")
    p = plan_info([one_1080p_420], copies=200)
    assert (p["sub_bits"], p["lookback_bits"], p["write_parts"], p["groups"], p["seg_bits"]) == (16384, 300, 4, 2, 2048)
