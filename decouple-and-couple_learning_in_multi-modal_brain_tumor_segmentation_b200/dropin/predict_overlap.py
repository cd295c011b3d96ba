"""Drop-in for the reference ``predict_overlap.py``: same ``tailor_and_concat`` /
``validate_softmax`` signatures, bodies routed through the CUDA library (dcl_b200.Engine).

Reference: predict_overlap.py:31-58 (8-corner tiling + crop-and-overwrite stitch, including its
5-voxel z shift) and :103-171 (evaluation loop: arg-max labels, label histogram, WT/TC/ET Dice).
"""
import numpy as np
import torch

from dcl_b200 import StitchMode, volio
from dcl_b200.engine import dice_from_counts


def _unwrap(model):
    return model.module if hasattr(model, "module") else model     # nn.DataParallel (test_overlap.py:78)


def _keep_scales(net, device, n=8):
    # one RNG draw per patch forward, in the reference's patch order
    return torch.cat([net.draw_keep_scale(1, device) for _ in range(n)], 0).numpy()


def tailor_and_concat(x, missing_modal, model, target=None):
    """(1,4,240,240,>=155) CUDA volume -> (1,4,240,240,155) stitched class probabilities."""
    net = _unwrap(model)
    eng = net.engine(x.device)
    with torch.cuda.device(x.device):
        out = eng.predict_volume(x.float(), StitchMode.REFERENCE, keep_scales=_keep_scales(net, x.device),
                                 want_probs=True, want_labels=False)
    return out["probs"]


def validate_softmax(valid_loader, model, load_file, multimodel, savepath='', names=None, verbose=False,
                     use_TTA=False, save_format=None, snapshot=False, visual='', postprocess=False,
                     valid_in_train=False):
    net = _unwrap(model)
    model.eval()
    wt_dices, tc_dices, et_dices = [], [], []
    print('sum=====', sum(p.numel() for p in model.parameters()))
    for i, data in enumerate(valid_loader):
        print('-------------------------------------------------------------------')
        msg = 'Subject {}/{}, '.format(i + 1, len(valid_loader))
        x, target = data[0], (data[1] if valid_in_train else None)
        x = x.cuda(non_blocking=True)
        eng = net.engine(x.device)
        tgt = None
        if target is not None:
            tgt = target[0, :, :, :155].to(x.device)
            tgt = torch.where(tgt == 4, torch.full_like(tgt, 3), tgt).to(torch.uint8)     # :150-152
        out = eng.predict_volume(x.float(), StitchMode.REFERENCE, keep_scales=_keep_scales(net, x.device),
                                 target=tgt, want_probs=False, want_labels=True)
        counts = out["counts"].cpu().numpy()
        soft = dice_from_counts(counts) if tgt is not None else [float('nan')] * 3
        name = names[i] if names is not None else str(i)
        print('name:{}, msg={}, DICE= WT:{},TC:{},ET:{}'.format(name, msg, soft[0], soft[1], soft[2]))
        print('0标签:{},1标签:{},2标签:{},3标签:{},索引最大值: {}'.format(
            counts[0], counts[1], counts[2], counts[3], int(np.max(np.nonzero(counts[:4])[0]))))
        wt_dices.append(soft[0]); tc_dices.append(soft[1]); et_dices.append(soft[2])
        if savepath:
            # the export block the reference keeps (commented out) in predict.py:310-350: '.npy' for ensembling,
            # '.nii.gz' with label 3 -> 4 for submission, optional per-frame snapshots
            volio.save_prediction(out["labels"], savepath, name, save_format or 'nii', snapshot=snapshot,
                                  visual=visual, verbose=verbose)
    print('WT Dice: %.4f' % np.mean(wt_dices))
    print('TC Dice: %.4f' % np.mean(tc_dices))
    print('ET Dice: %.4f' % np.mean(et_dices))
    return np.mean(wt_dices), np.mean(tc_dices), np.mean(et_dices)
