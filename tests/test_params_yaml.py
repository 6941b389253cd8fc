"""loadParamsYaml / `gbp_plan --params`: the reference's node is configured by `rosparam load config/params.yaml` plus the
launch files' state_publisher/* values (global_body_planner.cpp:15-28, :181-204, :220-228); the ROS-free driver reads the
same file and the same names.  Runs without a device (`--print-params` exits before any terrain is created)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_YAML = "/root/reference/config/params.yaml"

YAML = """\
topics:
  terrain_map: /terrain_map   # ignored by the planner's parameters
global_body_planner:
  update_rate: 1
  num_calls: 25            # comment after a value
  replan_time_limit: 0.75
  algorithm: "rrt-star-connect"
  state_action_pair_check_adaptive_step_size_flag: true
  cost_add_yaw:
    flag: true
    length_weight: 2
    yaw_weight: 0.5
  state_direction_sampling:
    flag: True
    probability_threshold: 0.3
    speed_direction_flag: true
  action_direction_sampling:
    flag: true
    probability_threshold: 1e-1

state_publisher:
  start_position_x: -1.5
  start_yaw: 0.25
  goal_position_x: 6
  goal_position_y: '2.5'
grid_map_visualization:
  grid_map_visualizations:
    - name: elevation_points
      type: point_cloud
      params:
        layer: elevation
"""


@pytest.fixture(scope="module")
def cli():
    from global_body_planner_b200 import build as b
    b.build()
    b.build_host()
    return b.CLI


def resolved(cli, *args):
    out = subprocess.run([cli, "default", *args, "--print-params"], check=True, capture_output=True, text=True).stdout
    return dict(line.rsplit(" ", 1) for line in out.strip().splitlines())


def test_defaults_are_the_params_yaml_defaults(cli):
    r = resolved(cli)
    assert r["global_body_planner/algorithm"] == "rrt-connect" and r["global_body_planner/num_calls"] == "1"
    assert float(r["global_body_planner/action_direction_sampling/probability_threshold"]) == 0.1  # config/params.yaml:27
    assert float(r["global_body_planner/state_direction_sampling/probability_threshold"]) == 0.05  # config/params.yaml:23
    assert float(r["state_publisher/goal_position_x"]) == 8.0  # launch/example.launch: (0, 0) -> (8, 0)


def test_yaml_overrides_and_later_options_win(cli, tmp_path):
    f = tmp_path / "params.yaml"
    f.write_text(YAML)
    r = resolved(cli, "--params", str(f))
    want = {"global_body_planner/num_calls": 25, "global_body_planner/replan_time_limit": 0.75,
            "global_body_planner/state_action_pair_check_adaptive_step_size_flag": 1, "global_body_planner/cost_add_yaw/flag": 1,
            "global_body_planner/cost_add_yaw/length_weight": 2, "global_body_planner/cost_add_yaw/yaw_weight": 0.5,
            "global_body_planner/state_direction_sampling/flag": 1, "global_body_planner/state_direction_sampling/probability_threshold": 0.3,
            "global_body_planner/state_direction_sampling/speed_direction_flag": 1, "global_body_planner/action_direction_sampling/flag": 1,
            "global_body_planner/action_direction_sampling/probability_threshold": 0.1, "state_publisher/start_position_x": -1.5,
            "state_publisher/start_position_y": 0, "state_publisher/start_yaw": 0.25, "state_publisher/goal_position_x": 6,
            "state_publisher/goal_position_y": 2.5, "state_publisher/goal_yaw": 0}
    assert r["global_body_planner/algorithm"] == "rrt-star-connect"
    for k, v in want.items():
        assert float(r[k]) == float(v), (k, r[k], v)
    r = resolved(cli, "--params", str(f), "--num-calls", "3", "--algorithm", "rrt-connect", "--action-direction-sampling", "0.4")
    assert r["global_body_planner/num_calls"] == "3" and r["global_body_planner/algorithm"] == "rrt-connect"
    assert float(r["global_body_planner/action_direction_sampling/probability_threshold"]) == 0.4
    r = resolved(cli, "--num-calls", "3", "--params", str(f))  # options apply in order
    assert r["global_body_planner/num_calls"] == "25"


@pytest.mark.parametrize("bad", ["global_body_planner:\n  num_calls: many\n", "global_body_planner:\n  cost_add_yaw:\n    flag: maybe\n",
                                 "global_body_planner:\n  algorithm: prm\n"])
def test_malformed_values_fail_loudly(cli, tmp_path, bad):
    f = tmp_path / "bad.yaml"
    f.write_text(bad)
    r = subprocess.run([cli, "default", "--params", str(f), "--print-params"], capture_output=True, text=True)
    assert r.returncode == 2 and "loadParamsYaml" in r.stderr
    r = subprocess.run([cli, "default", "--params", str(tmp_path / "absent.yaml"), "--print-params"], capture_output=True, text=True)
    assert r.returncode == 2 and "cannot open" in r.stderr


@pytest.mark.skipif(not os.path.exists(REF_YAML), reason="no /root/reference on this box")
def test_the_reference_s_own_params_yaml(cli):
    r = resolved(cli, "--params", REF_YAML)
    assert r["global_body_planner/num_calls"] == "1000" and float(r["global_body_planner/replan_time_limit"]) == 1.0
    assert r["global_body_planner/algorithm"] == "rrt-connect"
    for k in ("state_action_pair_check_adaptive_step_size_flag", "cost_add_yaw/flag", "state_direction_sampling/flag", "action_direction_sampling/flag"):
        assert r["global_body_planner/" + k] == "0"
    assert float(r["global_body_planner/action_direction_sampling/probability_threshold"]) == 0.1
    assert float(r["global_body_planner/state_direction_sampling/probability_threshold"]) == 0.05
