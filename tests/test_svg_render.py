"""render_svg (host-side SVG animation, no GPU): well-formed XML, one animated circle per agent, finished agents
fade out, static mode has no <animate>."""
import xml.etree.ElementTree as ET

import numpy as np

from pogema_b200.wrappers import AgentState, AnimationConfig, render_svg

NS = "{http://www.w3.org/2000/svg}"


def _history():
    a0 = [AgentState(0, 0, 3, 3, 0, True), AgentState(0, 1, 3, 3, 1, True), AgentState(1, 1, 3, 3, 2, True)]
    a1 = [AgentState(3, 0, 3, 1, 0, True), AgentState(3, 1, 3, 1, 1, False), AgentState(3, 1, 3, 1, 2, False)]
    return [a0, a1]


def test_animated_svg_structure():
    ob = np.zeros((4, 4), np.uint8)
    ob[1, 2] = ob[2, 2] = 1
    root = ET.fromstring(render_svg(ob, _history(), AnimationConfig(), obs_radius=2))
    agents = [c for c in root.iter(NS + "circle") if c.get("class") == "a"]
    assert len(agents) == 2
    anims = [{a.get("attributeName"): a.get("values") for a in c.iter(NS + "animate")} for c in agents]
    assert anims[0]["cx"] == "150;250;250" and anims[0]["cy"] == "150;150;250"   # x = row -> svg y, y = column -> svg x
    assert "opacity" not in anims[0] and anims[1]["opacity"] == "1.0;0.0;0.0"      # the finished agent disappears
    assert len(list(root.iter(NS + "rect"))) == 2 + 4 * 4 + 4                      # obstacles + wall ring


def test_static_and_egocentric_variants():
    ob = np.zeros((4, 4), np.uint8)
    ob[0, 3] = 1
    static = ET.fromstring(render_svg(ob, _history(), AnimationConfig(static=True, show_lines=True, show_border=False), 2))
    assert not list(static.iter(NS + "animate")) and len(list(static.iter(NS + "line"))) == 2
    assert len(list(static.iter(NS + "rect"))) == 1
    ego = ET.fromstring(render_svg(ob, _history(), AnimationConfig(egocentric_idx=0), 1))
    fills = [c.get("fill") for c in ego.iter(NS + "circle") if c.get("class") == "a"]
    assert fills[0] != fills[1]
    no_agents = ET.fromstring(render_svg(ob, _history(), AnimationConfig(show_agents=False), 1))
    assert not list(no_agents.iter(NS + "circle"))
