"""SimulatorTrajectory behaviour, mirroring the reference's mythos/simulators/tests/test_io.py (slice, filter,
with_state_metadata, addition / concat with metadata merge, box sizes, temperatures) on torch tensors."""

import pytest
import torch

from mythos_b200.rigid_body import Quaternion
from mythos_b200.simulators.io import SimulatorTrajectory


def traj(n, fill=0.0, **kw):
    return SimulatorTrajectory(center=torch.full((n, 2, 3), fill), orientation=Quaternion(torch.full((n, 2, 4), fill)), **kw)


def test_slice_slices_every_per_state_field():
    t = traj(6, box_size=torch.arange(18.0).reshape(6, 3), temperature=torch.arange(6.0), metadata={"v": torch.arange(6)})
    s = t.slice(slice(2, 5))
    assert s.length() == 3 and s.box_size.shape == (3, 3) and s.box_size[0, 0] == 6.0
    assert s.temperature.tolist() == [2.0, 3.0, 4.0] and s.metadata["v"].tolist() == [2, 3, 4]
    assert t.slice(1).length() == 1
    assert t.slice([0, 5]).metadata["v"].tolist() == [0, 5]
    assert traj(3).slice(slice(0, 2)).temperature is None


def test_with_state_metadata_and_filter():
    t = traj(5).with_state_metadata(force=10.0, torque=5.0)
    assert all(v.shape[0] == 5 for v in t.metadata.values()) and t.metadata["force"][3] == 10.0
    t = SimulatorTrajectory(center=torch.arange(30.0).reshape(10, 1, 3), orientation=Quaternion(torch.zeros(10, 1, 4)),
                            metadata={"value": torch.arange(10)})
    f = t.filter(lambda md: md["value"] % 2 == 0)
    assert f.length() == 5 and torch.all(f.metadata["value"] % 2 == 0) and f.center[1, 0, 0] == 6.0
    assert t.filter(lambda md: md["value"] > 100).length() == 0


def test_addition_merges_metadata_and_box_sizes():
    a = traj(3, 1.0, box_size=torch.ones(3, 3) * 10, metadata={"value": torch.tensor([1, 1, 1]), "value2": torch.tensor([2, 2, 2]),
                                                               "value3": torch.full((3, 2), 3), "value4": torch.full((3, 2), 4)})
    b = traj(2, 0.0, box_size=torch.zeros(2, 3), metadata={"value": torch.tensor([0, 0]), "value3": torch.zeros(2, 2, dtype=torch.long)})
    c = a + b
    assert c.length() == 5 and all(v.shape[0] == 5 for v in c.metadata.values())
    assert c.metadata["value"].tolist() == [1, 1, 1, 0, 0]
    assert c.metadata["value2"][:3].tolist() == [2, 2, 2] and torch.all(torch.isnan(c.metadata["value2"][3:]))
    assert c.metadata["value3"].shape == (5, 2) and torch.all(torch.isnan(c.metadata["value4"][3:]))
    assert c.box_size.shape == (5, 3) and c.box_size[0].tolist() == [10, 10, 10] and c.box_size[3].tolist() == [0, 0, 0]
    assert (traj(2) + traj(3)).metadata is None and (traj(2) + traj(3)).box_size is None


def test_concat_errors_and_temperatures():
    with pytest.raises(ValueError, match="empty list"):
        SimulatorTrajectory.concat([])
    one = traj(2)
    assert SimulatorTrajectory.concat([one]) is one
    with pytest.raises(ValueError, match="box sizes"):
        traj(2, box_size=torch.ones(2, 3)) + traj(2)
    with pytest.raises(ValueError, match="temperatures"):
        traj(2, temperature=torch.ones(2)) + traj(2)
    with pytest.raises(ValueError, match="mismatched shapes"):
        traj(2, metadata={"v": torch.zeros(2, 3)}) + traj(2, metadata={"v": torch.zeros(2, 4)})
    c = SimulatorTrajectory.concat([traj(2, temperature=torch.full((2,), 0.1)), traj(3, temperature=torch.full((3,), 0.2)), traj(1, temperature=torch.ones(1))])
    assert c.temperature.tolist() == pytest.approx([0.1, 0.1, 0.2, 0.2, 0.2, 1.0])


def test_sharded_blocks_cannot_be_resliced_or_concatenated():
    blk = traj(4, temperature=torch.ones(4), shard=(4, 8, 16))
    assert blk.slice(slice(0, 4)).shard == (4, 8, 16)  # whole block: bounds still valid
    with pytest.raises(ValueError, match="frame-sharded"):
        blk.slice(slice(1, 4))
    with pytest.raises(ValueError, match="frame-sharded"):
        blk + blk
