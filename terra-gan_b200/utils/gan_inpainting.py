"""Drop-in for the reference's utils/gan_inpainting.py:5-19 (`inpaint_with_gan`), the inference path the north star
names: same signature and output file, running on the B200 path through mvp_gan.src.evaluate."""
import logging
from pathlib import Path

from mvp_gan.src.evaluate import evaluate, evaluate_batch


def inpaint_with_gan(dem_image_path, mask_path, output_dir, checkpoint_path):
    output_dir = Path(output_dir)
    dem_image_path = Path(dem_image_path)
    output_dir.mkdir(parents=True, exist_ok=True)
    inpainted_image_path = output_dir / f"{dem_image_path.stem}_inpainted.png"
    evaluate(dem_image_path, mask_path, checkpoint_path, inpainted_image_path)
    logging.info(f"Inpainted image saved to {inpainted_image_path}")
    return inpainted_image_path


def inpaint_many_with_gan(dem_image_paths, mask_paths, output_dir, checkpoint_path, batch: int = 16):
    """Batched form: one generator forward per `batch` tiles."""
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    outs = [output_dir / f"{Path(p).stem}_inpainted.png" for p in dem_image_paths]
    evaluate_batch(dem_image_paths, mask_paths, checkpoint_path, outs, batch=batch)
    return outs
