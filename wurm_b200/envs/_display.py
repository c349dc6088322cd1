"""Host-side display of a batch of int16 (N,3,S,S) images: what the reference's `render()` methods do after they
have the image (wurm/envs/single_snake.py:389-428, multi_snake.py:229-266).  Outside the hot path; kept so that
`env.render('rgb_array')` keeps working for callers of the reference's API."""
import numpy as np


def show(env, images, mode: str = 'human', index: int = None):
    if mode not in ('human', 'rgb_array'):
        raise ValueError('Render mode not recognised.')
    from PIL import Image
    rows, cols = (1, 1) if (env.num_envs == 1 or index is not None) else (env.render_args['num_rows'], env.render_args['num_cols'])
    first = index or 0
    tiles = images[first:first + rows * cols].permute(0, 2, 3, 1)                 # (rows*cols, S, S, 3)
    S = tiles.shape[1]
    mosaic = tiles.reshape(rows, cols, S, S, 3).permute(0, 2, 1, 3, 4).reshape(rows * S, cols * S, 3)
    side = env.render_args['size']
    frame = np.asarray(Image.fromarray(mosaic.cpu().numpy().astype(np.uint8)).resize((side * cols, side * rows)))
    if mode == 'rgb_array':
        return frame
    if env.viewer is None:
        from gym.envs.classic_control import rendering
        env.viewer = rendering.SimpleImageViewer()
    env.viewer.imshow(frame)
    return env.viewer.isopen
