"""VisualCLAProcessor: tokenizer + CLIP image processor bundle.  The reference exports it
(ref: models/visualcla/processing_visualcla.py:11-131) but none of its scripts use it; kept as a thin
convenience so `from visualcla import VisualCLAProcessor` keeps working."""


class VisualCLAProcessor:
    attributes = ["image_processor", "tokenizer"]

    def __init__(self, image_processor=None, tokenizer=None, **kwargs):
        if image_processor is None:
            raise ValueError("You need to specify an `image_processor`.")
        if tokenizer is None:
            raise ValueError("You need to specify a `tokenizer`.")
        self.image_processor, self.tokenizer = image_processor, tokenizer

    def __call__(self, text=None, images=None, return_tensors=None, **kwargs):
        if text is None and images is None:
            raise ValueError("You have to specify either text or images. Both cannot be none.")
        enc = self.tokenizer(text, return_tensors=return_tensors, **kwargs) if text is not None else None
        img = self.image_processor(images, return_tensors=return_tensors, **kwargs) if images is not None else None
        if enc is not None and img is not None:
            enc["pixel_values"] = img.pixel_values
            return enc
        return enc if enc is not None else img

    def batch_decode(self, *args, **kwargs):
        return self.tokenizer.batch_decode(*args, **kwargs)

    def decode(self, *args, **kwargs):
        return self.tokenizer.decode(*args, **kwargs)
