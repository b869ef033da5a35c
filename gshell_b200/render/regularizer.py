"""Image-space regularisers of tick() behind the names of the reference's render/regularizer.py (:21-52), evaluated by the
fused reduction kernels of csrc/tick_ops.cu (gshell_b200/losses.py); tick() itself asks for all of its terms in ONE pass."""
from .. import losses


def chroma_loss(kd, color_ref, lambda_chroma):
    return losses.image_terms(color_ref, losses.T_CHROMA, (lambda_chroma, 0, 0, 0, 0, 0), kd=kd)[1]


def shading_loss(diffuse_light, specular_light, color_ref, lambda_diffuse, lambda_specular):
    return losses.image_terms(color_ref, losses.T_SHADING, (0, lambda_diffuse, lambda_specular, 0, 0, 0), diffuse=diffuse_light,
                              specular=specular_light)[1]


def material_smoothness_grad(kd_grad, ks_grad, nrm_grad, lambda_kd=0.25, lambda_ks=0.1, lambda_nrm=0.0):
    return losses.image_terms(kd_grad.new_zeros(kd_grad.shape[:-1] + (4,)), losses.T_SMOOTH, (0, 0, 0, lambda_kd, lambda_ks, lambda_nrm),
                              kd_grad=kd_grad, ks_grad=ks_grad, nrm_grad=nrm_grad)[1]
