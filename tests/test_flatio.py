"""-m "not gpu" / gpu: the flat-file stand-in of the reference's HDF5 I/O (bioslam_b200/flatio.py): APDM-v5 dataset
names in, saveResultsToH5File dataset names out."""
import os
import numpy as np
import pytest
from bioslam_b200 import flatio, synth


def test_apdm_trial_round_trip(tmp_path):
    tr = synth.make_trial(seed=2, duration_s=2.0)
    f = str(tmp_path / "trial.npz")
    flatio.write_apdm_trial(f, tr)
    assert flatio.is_apdm_v5(f)
    m = flatio.read_imu_map(f)
    assert sorted(m) == sorted(flatio.LOWER_BODY_LABELS)                       # looked up by 'Label 0' like the reference
    for i, label in enumerate(flatio.LOWER_BODY_LABELS):
        s = m[label]
        assert np.array_equal(s["acc"], tr.acc[i]) and np.array_equal(s["gyro"], tr.gyro[i])
        assert s["time"][0] == 0.0 and np.abs(np.diff(s["time"]) - tr.dt).max() < 1e-6      # unix microseconds -> seconds from the first sample
        assert s["quat"].shape == (tr.acc.shape[1], 4) and s["id"] == 1000 + i
    # not an APDM v5 file -> the reference throws
    np.savez(str(tmp_path / "other.npz"), x=np.zeros(3))
    with pytest.raises(RuntimeError):
        flatio.read_imu_map(str(tmp_path / "other.npz"))


def test_quaternion_of_rotation():
    R = synth.exp_so3(np.random.default_rng(0).normal(size=(50, 3)) * 2.0)
    q = flatio._rot_to_quat(R)
    assert np.abs(np.linalg.norm(q, axis=1) - 1.0).max() < 1e-12 and np.all(q[:, 0] >= 0)
    w, x, y, z = q.T
    R2 = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                   np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                   np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)
    assert np.abs(R2 - R).max() < 1e-12


@pytest.mark.gpu
def test_results_file_has_the_reference_dataset_names(tmp_path):
    from bioslam_b200 import host
    tr = synth.make_trial(seed=8, duration_s=3.0)
    f = str(tmp_path / "trial.npz")
    flatio.write_apdm_trial(f, tr)
    m = flatio.read_imu_map(f)
    # the estimators are built from the FILE (exampleLowerBodyEstimation.cpp:14-66)
    from types import SimpleNamespace
    from_file = SimpleNamespace(acc=np.stack([m[l]["acc"] for l in flatio.LOWER_BODY_LABELS]), gyro=np.stack([m[l]["gyro"] for l in flatio.LOWER_BODY_LABELS]),
                                time=m["Sacrum"]["time"])
    lb = host.lower_body_from_trial(from_file)
    lb.set(m_maxIterations=8, m_maxNumRestarts=0)
    lb.setup(); lb.fastOptimize()
    out = str(tmp_path / "results.npz")
    flatio.write_results(out, lb)
    z = np.load(out)
    K = int(z["/General/nKeyframes"])
    for name in ("/Estimated/R_SacrumImu_to_N", "/Estimated/q_LFootImu_to_N", "/Estimated/p_RThighImu", "/Estimated/v_LShankImu", "/Estimated/gyrobias_RFootImu",
                 "/Estimated/accelbias_SacrumImu", "/Estimated/RShankImuAngVel", "/Estimated/SacrumImuToLHipCtr", "/Estimated/RKneeAxis_RShank",
                 "/Estimated/LAnkleAxis_Foot", "/Derived/RKneeAngles_FlexEx", "/Derived/LHipAngles_IntExtRot"):
        assert name in z.files, name
    assert z["/Estimated/R_SacrumImu_to_N"].shape == (K, 3, 3) and z["/Estimated/p_RThighImu"].shape == (K, 3) and z["/Derived/RKneeAngles_FlexEx"].shape == (K,)
    assert abs(np.linalg.norm(z["/Estimated/RKneeAxis_RShank"]) - 1.0) < 1e-12
