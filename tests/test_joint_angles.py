"""Derived joint angles.  The oracle restatement (oracle/jointangles.hpp) is pinned the way the
reference pins bioutils::consistent3DofJcsAngles (test/unit/testJointAngleCalculation.cpp): zero angles for identical
segments (:140-152), a spin of the distal IMU about the proximal flexion axis / its own long axis / the floating axis
adds exactly that angle to one joint angle and leaves the other two alone (:24-136), tolerance 1e-10 rad as upstream.
The CUDA kernel (b200_lower_body_joint_angles) is checked against the oracle under the emulator here and on the GPU in
tests/test_gpu_parity.py."""
import os
import subprocess
import numpy as np
import pytest
from bioslam_b200 import abi, graphs, lib, synth
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "cudaemu", "libbioslam_b200_emu.so")


def rand_rot(rng):
    return po.so3_expmap(rng.normal(size=3) * 1.3)


def axis_angle(axis, ang):
    axis = np.asarray(axis, dtype=float) / np.linalg.norm(axis)
    return po.so3_expmap(axis * ang)


def wrap_pi(a):
    return np.remainder(a + np.pi, 2 * np.pi) - np.pi


def unit(v):
    return np.asarray(v, dtype=float) / np.linalg.norm(v)


def test_zero_angles_for_identical_segments(oracle_lib):
    rng = np.random.default_rng(0)
    for _ in range(200):
        R = rand_rot(rng)
        assert np.abs(po.jcs_angles(R, R)).max() < 1e-10


@pytest.mark.parametrize("random_frames,same_imu", [(True, False), (True, True), (False, True), (False, False)])
def test_expected_joint_angle_response(oracle_lib, random_frames, same_imu):
    rng = np.random.default_rng(1)
    for _ in range(60):
        Rp_imu = rand_rot(rng)
        if random_frames:
            p_right = unit(rng.normal(size=3)); p_prox = np.cross(unit(rng.normal(size=3)), p_right)
            d_right = unit(rng.normal(size=3)); d_prox = np.cross(unit(rng.normal(size=3)), d_right)
        else:
            p_right = np.array([0, 0, 1.0]); p_prox = np.array([-1.0, 0, 0]); d_right = p_right.copy(); d_prox = p_prox.copy()
        Rd_imu = Rp_imu if same_imu else rand_rot(rng)
        Sp = po.R_imu_to_segment(p_right, p_prox, -p_prox)
        Sd = po.R_imu_to_segment(d_right, d_prox, -d_prox)
        assert abs(np.linalg.det(Sp) - 1) < 1e-10 and np.abs(Sp @ Sp.T - np.eye(3)).max() < 1e-10
        Rp = po.R_segment_to_N(Rp_imu, Sp)
        a0 = po.jcs_angles(Rp, po.R_segment_to_N(Rd_imu, Sd))
        flex_axis_in_distal = Rd_imu.T @ Rp_imu @ p_right
        for _ in range(5):
            d = rng.uniform(-0.4, 0.4)
            # flexion: rotate the distal IMU by -d about the proximal flexion axis (expressed in the distal IMU frame)
            a1 = po.jcs_angles(Rp, po.R_segment_to_N(Rd_imu @ axis_angle(flex_axis_in_distal, -d).T, Sd))
            assert abs(wrap_pi(a0[0] + d) - a1[0]) < 1e-10 and abs(a0[1] - a1[1]) < 1e-10 and abs(a0[2] - a1[2]) < 1e-10
            # internal/external rotation: about the distal segment's own proximal vector
            a2 = po.jcs_angles(Rp, po.R_segment_to_N(Rd_imu @ axis_angle(d_prox, -d).T, Sd))
            assert abs(wrap_pi(a0[2] + d) - a2[2]) < 1e-10 and abs(a0[0] - a2[0]) < 1e-10 and abs(a0[1] - a2[1]) < 1e-10
            # ab/adduction: about the floating axis cross(k, I)
            spin = unit(np.cross(d_prox, flex_axis_in_distal))
            a3 = po.jcs_angles(Rp, po.R_segment_to_N(Rd_imu @ axis_angle(spin, d).T, Sd))
            exp = wrap_pi(a0[1] + d)
            if abs(exp) <= np.pi / 2:
                assert abs(exp - a3[1]) < 1e-10 and abs(a0[0] - a3[0]) < 1e-10 and abs(a0[2] - a3[2]) < 1e-10


def test_clinical_angles_mirror_the_left_leg(oracle_lib):
    rng = np.random.default_rng(2)
    Rp, Rd = rand_rot(rng), rand_rot(rng)
    c = po.jcs_angles(Rp, Rd)
    assert np.allclose(po.jcs_angles(Rp, Rd, clinical=True, right_leg=True), c)
    assert np.allclose(po.jcs_angles(Rp, Rd, clinical=True, right_leg=False), c * [1, -1, -1])


def test_pelvis_frame(oracle_lib):
    R = po.R_sacrum_to_pelvis([-1, 0, 0], [0.0, 0.1, 0.09], [0.0, 0.1, -0.09])
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(R) - 1) < 1e-12
    assert np.allclose(R[0], unit(np.array([0, 0, 0.18])))      # x = hip-centre line (master vector)
    assert R[2] @ np.array([-1.0, 0, 0]) > 0.99                 # z close to the supplied up vector


def test_truth_angles_of_the_synthetic_gait(oracle_lib):
    """the generator's knees are pure hinges: the oracle angles computed from TRUE poses and TRUE statics must show
    knee flexion with (almost) no ab/adduction or axial rotation"""
    tr = synth.make_trial(seed=2, duration_s=3.0)
    g, tv, sv = graphs.lower_body_graph(tr, preint=po.preintegrate)
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv, rot_noise=0.0, pos_noise=0.0)
    soff = g.static_value_offsets()
    sidx = {n: i for i, n in enumerate(graphs.STATIC_POINTS + graphs.STATIC_AXES)}
    for n in graphs.STATIC_POINTS:
        sv[soff[sidx[n]]:soff[sidx[n]] + 3] = tr.true_offsets[n]
    for n in graphs.STATIC_AXES:
        sv[soff[sidx[n]]:soff[sidx[n]] + 3] = unit(tr.true_axes[n])
    ang = po.lower_body_joint_angles(g, tv, sv, graphs.lower_body_joint_angle_desc())
    assert np.all(np.isfinite(ang))
    rknee = np.degrees(ang[:, 0:3])
    assert np.ptp(rknee[:, 0]) > 20.0                          # the knee flexes through tens of degrees
    assert np.ptp(rknee[:, 1]) < 2.0 and np.ptp(rknee[:, 2]) < 2.0


def test_kernel_matches_oracle_under_emulator(oracle_lib):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cudaemu")])
    tr = synth.make_trial(seed=6, duration_s=1.5)
    g, tv, sv = graphs.lower_body_graph(tr, preint=po.preintegrate)
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    d = graphs.lower_body_joint_angle_desc()
    p = lib.Problem(g, so_path=EMU, chunks=1); p.set_values(tv, sv)
    a, b = p.joint_angles(d), po.lower_body_joint_angles(g, tv, sv, d)
    assert a.shape == (g.K, 12) and np.abs(a - b).max() < 1e-12
    bad = abi.JointAngleDesc(); bad.pose_var[0] = 1   # a velocity variable
    with pytest.raises(lib.B200Error):
        p.joint_angles(bad)
