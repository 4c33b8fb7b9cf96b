"""Golden vectors for the RE10K chunk format from the REFERENCE's own code: `DatasetRE10kStyle.convert_poses` /
`convert_images` (src/dataset/dataset_re10k_style.py:215-246) and `camera_normalization` (src/misc/cam_utils.py:27-42),
run on a small synthetic chunk entry (3 JPEG frames + camera rows).  The two methods are executed from the reference
source as plain functions (the dataset class itself needs the hydra / lightning stack).
python tests/golden/make_chunk_golden.py"""
import ast
import sys
import textwrap
from io import BytesIO
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/src")


def ref_functions():
    src = (REF / "dataset/dataset_re10k_style.py").read_text()
    tree = ast.parse(src)
    ns = {}
    import torchvision.transforms as tf
    from einops import rearrange, repeat
    from PIL import Image
    ns.update(torch=torch, rearrange=rearrange, repeat=repeat, Image=Image, BytesIO=BytesIO, Float=None, UInt8=None, Tensor=torch.Tensor)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ("convert_poses", "convert_images"):
            node.returns = None
            for a in node.args.args:
                a.annotation = None
            code = ast.unparse(node)
            exec(textwrap.dedent(code), ns)
    cam = {}
    exec((REF / "misc/cam_utils.py").read_text().split("####### Pose update from delta")[0], cam)

    class Self:
        to_tensor = tf.ToTensor()
    return (lambda p: ns["convert_poses"](Self(), p)), (lambda im: ns["convert_images"](Self(), im)), cam["camera_normalization"]


def main():
    from PIL import Image
    convert_poses, convert_images, camera_normalization = ref_functions()
    g = torch.Generator().manual_seed(11)
    n = 3
    # camera rows: fx fy cx cy 0 0 + w2c 3x4
    from scipy.spatial.transform import Rotation as R
    rows = []
    for i in range(n):
        w2c = np.eye(4)
        w2c[:3, :3] = R.from_euler("xyz", [0.02 * i, -0.1 * i, 0.01]).as_matrix()
        w2c[:3, 3] = [0.3 * i, -0.02 * i, 0.1 * i]
        rows.append(np.concatenate([[0.52, 0.93, 0.5, 0.5, 0, 0], w2c[:3].reshape(-1)]))
    cameras = torch.tensor(np.stack(rows), dtype=torch.float32)
    images = []
    for i in range(n):
        yy, xx = np.mgrid[0:36, 0:64]
        arr = np.stack([(127 + 120 * np.sin(xx / 7.0 + i + c) * np.cos(yy / 5.0)).astype(np.uint8) for c in range(3)], -1)
        buf = BytesIO()
        Image.fromarray(arr).save(buf, format="JPEG", quality=90)
        images.append(torch.tensor(np.frombuffer(buf.getvalue(), dtype=np.uint8).copy()))
    extr, intr = convert_poses(cameras)
    imgs = convert_images(images)
    norm = camera_normalization(extr[0:1], extr)
    out = dict(cameras=cameras.numpy(), extrinsics=extr.numpy(), intrinsics=intr.numpy(), images=imgs.numpy(),
               normalized=norm.numpy(), n_jpeg=np.array([len(j) for j in images]),
               jpeg=np.concatenate([j.numpy() for j in images]))
    dst = Path(__file__).parent / "chunk_golden.npz"
    np.savez_compressed(dst, **out)
    print(dst, imgs.shape, extr.shape)


if __name__ == "__main__":
    main()
