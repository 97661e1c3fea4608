"""CPU test of the host-side bookkeeping of pipeline.DetectPipeline (no GPU): the loop's order of operations - upload batch
i+1, queue batch i, hand batch i - slots to the caller - replayed on the resource indices `pipeline.schedule` assigns, checking
that no device input buffer and no pinned result slot is re-filled while its previous user still needs it, and that a
pipeline object refuses to exist without a CUDA device (there is no CPU path)."""
import pytest
import torch

from yolov5_obb_b200.pipeline import DetectPipeline, schedule


@pytest.mark.parametrize("slots", [1, 2, 3, 4])
def test_schedule_never_reuses_a_busy_resource(slots):
    n_in = max(2, slots)
    n_batches = 23
    queued, handed = set(), set()
    in_owner, host_owner = {}, {}          # resource index -> batch that filled it last
    pending = []

    def upload(b):
        _, j, _, _ = schedule(b, slots, n_in)
        prev = in_owner.get(j)
        # the copy stream waits for the EVENT of the previous user's layout pass: that event exists only once it was queued
        assert prev is None or prev in queued, (b, j, prev)
        in_owner[j] = b

    upload(0)
    for i in range(n_batches):
        slot, j, jn, hs = schedule(i, slots, n_in)
        assert 0 <= slot < slots and 0 <= j < n_in and 0 <= hs <= slots
        assert jn == schedule(i + 1, slots, n_in)[1]
        if i + 1 < n_batches:
            upload(i + 1)
        assert in_owner[j] == i            # nobody overwrote batch i's input before it was queued
        prev = host_owner.get(hs)
        assert prev is None or prev in handed, (i, hs, prev)   # the pinned slot's previous result was copied out already
        host_owner[hs] = i
        queued.add(i)
        pending.append(i)
        if len(pending) > slots:
            handed.add(pending.pop(0))
        # batches on the same slot run in stream order; at most `slots` batches are queued and not yet handed over ... + the one just queued
        assert len(pending) <= slots
    while pending:
        handed.add(pending.pop(0))
    assert handed == set(range(n_batches))


def test_same_slot_batches_are_stream_ordered():
    """two batches that share a compute slot (hence its plan's buffers and its post-process graph) are always i and i + k*slots:
    they are queued on ONE stream, in order - the only protection the slot's device buffers need"""
    for slots in (1, 2, 3):
        by_slot = {}
        for i in range(12):
            by_slot.setdefault(schedule(i, slots, max(2, slots))[0], []).append(i)
        for s, bs in by_slot.items():
            assert bs == sorted(bs) and all(b % slots == s for b in bs)


def test_pipeline_needs_a_cuda_device():
    m = torch.nn.Linear(2, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        DetectPipeline(m, device="cpu")
