"""``timm.models.vision_transformer.VisionTransformer(img_size, patch_size, in_chans, embed_dim, depth, num_heads,
num_classes=0)``: the headless ViT the reference builds, with timm's parameter names; ``forward_features`` runs on the
sm_100a kernels."""


def __getattr__(name):
    if name == "VisionTransformer":
        from ssrl_vit_mae_jepa_b200.classifier import VisionTransformer

        return VisionTransformer
    raise AttributeError(name)
