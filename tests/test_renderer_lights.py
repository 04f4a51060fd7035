"""CPU: how Panda3dLightData lists become the arrays hpb_render consumes (panda3d_batch_renderer.lights_from_light_datas)."""
import numpy as np
import pytest

from happypose_b200.renderer.panda3d_batch_renderer import ambient_from_light_datas, lights_from_light_datas, make_scene_lights
from happypose_b200.renderer.types import Panda3dLightData


def test_ambient_only_scenes():
    one = [Panda3dLightData(light_type="ambient", color=(1.0, 1.0, 1.0, 1.0))]
    assert ambient_from_light_datas([one, one]) is None                       # the hot path: nothing to upload
    three = 3 * [Panda3dLightData(light_type="ambient", color=(1.0, 1.0, 1.0, 1.0))]  # test_batch_renderer_panda3d.py:52-60
    amb = ambient_from_light_datas([one, three])
    np.testing.assert_array_equal(amb, [[1, 1, 1], [3, 3, 3]])
    dim = [Panda3dLightData(light_type="ambient", color=(0.7, 0.8, 0.9, 1.0))]
    np.testing.assert_allclose(ambient_from_light_datas([dim]), [[0.7, 0.8, 0.9]])


def test_scene_light_rig_positions_follow_the_bounding_radius():
    amb, lights = lights_from_light_datas([make_scene_lights(), make_scene_lights()], radii=[0.1, 0.25])
    np.testing.assert_allclose(amb, np.full((2, 3), 0.1))
    assert lights.shape == (2, 6, 8) and (lights[:, :, 0] == 0).all()         # six point lights
    np.testing.assert_allclose(lights[0, :, 1:4], np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]) * 1.0)
    np.testing.assert_allclose(lights[1, :, 1:4], np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]) * 2.5)
    np.testing.assert_allclose(lights[:, :, 4:7], 0.4)


def test_mixed_scenes_are_padded_and_directional_lights_keep_a_direction():
    def sun(root, node):
        node.setPos(0, 0, 2)
        node.lookAt(0, 0, 0)

    scenes = [[Panda3dLightData("ambient", (0.2, 0.2, 0.2, 1)), Panda3dLightData("directional", (0.5, 0.4, 0.3, 1), sun)],
              [Panda3dLightData("ambient")]]
    amb, lights = lights_from_light_datas(scenes, radii=[1.0, 1.0])
    assert lights.shape == (2, 1, 8)
    np.testing.assert_allclose(lights[0, 0], [1, 0, 0, -1, 0.5, 0.4, 0.3, 0])
    assert not lights[1].any()                                                   # padding = black light
    np.testing.assert_allclose(amb, [[0.2, 0.2, 0.2], [1, 1, 1]])
    with pytest.raises(NotImplementedError):
        lights_from_light_datas([[Panda3dLightData("spot", positioning_function=sun)]], [1.0])
    with pytest.raises(AssertionError):
        lights_from_light_datas([[Panda3dLightData("point")]], [1.0])            # no positioning function (scene renderer :303)
