"""Test-side wrappers of bioslam_b200.graphs that default the preintegration to the CPU oracle."""
from bioslam_b200.graphs import *  # noqa: F401,F403
from bioslam_b200 import graphs as _g


def _oracle_preint(combined):
    from oracle import pyoracle
    return lambda acc, gyro, dt, n_per, bh: pyoracle.preintegrate(acc, gyro, dt, n_per, bh, combined=combined)


def lower_body_graph(trial, n_per=10, preint=None, use_joint_vel=False, **kw):
    return _g.lower_body_graph(trial, n_per, preint or _oracle_preint(True), use_joint_vel, **kw)


def single_imu_graph(acc, gyro, dt, n_per=10, static_bias=False, preint=None, init_R=None, ref_local=(0.0, 1.0, 0.0)):
    return _g.single_imu_graph(acc, gyro, dt, n_per, static_bias, preint or _oracle_preint(not static_bias), init_R, ref_local)
