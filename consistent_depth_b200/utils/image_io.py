"""`.raw` float32 image I/O — the on-disk format of the reference's utils/image_io.py:101-169:
24-byte header (int32 h, int32 w, int32 cv_type, uint64 pixel_size) followed by HWC float32."""
import struct

import numpy as np

CV_32F, CV_CN_SHIFT = 5, 3


def load_raw_float32_image(file_name):
    with open(file_name, "rb") as f:
        h, w, cv_type = struct.unpack("iii", f.read(12))
        pixel_size = struct.unpack("Q", f.read(8))[0]
        d = ((cv_type - CV_32F) >> CV_CN_SHIFT) + 1
        if d < 1 or d != pixel_size // 4:
            raise Exception("Incompatible pixel_size(%d) and cv_type(%d)" % (pixel_size, cv_type))
        data = np.frombuffer(f.read(), dtype=np.float32)
    return data.reshape(h, w) if d == 1 else data.reshape(h, w, d)


def save_raw_float32_image(file_name, image):
    image = np.ascontiguousarray(image, dtype=np.float32)
    if image.ndim == 2:
        h, w = image.shape
        d = 1
    else:
        h, w, d = image.shape
    with open(file_name, "wb") as f:
        f.write(struct.pack("iii", h, w, CV_32F + ((d - 1) << CV_CN_SHIFT)))
        f.write(struct.pack("Q", d * 4))
        f.write(image.tobytes("C"))
