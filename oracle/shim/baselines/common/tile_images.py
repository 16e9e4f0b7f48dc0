def tile_images(img_nhwc):
    return img_nhwc
