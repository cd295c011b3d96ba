"""Drop-in for the reference ``predict_cls.py`` evaluation path with 8-flip test-time augmentation
(predict_cls.py:180-203): same ``tailor_and_concat`` / ``validate_softmax`` names, bodies routed through the CUDA
library.  ``use_TTA=True`` runs ``dcl_predict_volume_tta`` (flip -> tiling -> un-flip -> softmax -> mean over the 8
flips, all on the device); ``use_TTA=False`` is the plain ``predict_overlap`` path."""
import numpy as np
import torch

import predict_overlap as _po
from dcl_b200.engine import dice_from_counts

tailor_and_concat = _po.tailor_and_concat          # byte-identical in the reference (predict_cls.py:31-58)


def tta_tailor_and_concat(x, missing_modal, model):
    """(1,4,240,240,>=155) CUDA volume -> (1,4,240,240,155): mean over the 8 flips of softmax(tailor_and_concat)."""
    net = _po._unwrap(model)
    eng = net.engine(x.device)
    with torch.cuda.device(x.device):
        keeps = _po._keep_scales(net, x.device, 64).reshape(8, 8, 16)
        out = eng.predict_volume_tta(x.float(), keep_scales=keeps, want_probs=True, want_labels=False)
    return out["probs"]


def validate_softmax(valid_loader, model, load_file=None, multimodel=False, savepath='', names=None, verbose=False,
                     use_TTA=False, save_format=None, snapshot=False, visual='', postprocess=False,
                     valid_in_train=False):
    if not use_TTA:
        return _po.validate_softmax(valid_loader, model, load_file, multimodel, savepath, names, verbose, use_TTA,
                                    save_format, snapshot, visual, postprocess, valid_in_train)
    net = _po._unwrap(model)
    model.eval()
    wt_dices, tc_dices, et_dices = [], [], []
    for i, data in enumerate(valid_loader):
        msg = 'Subject {}/{}, '.format(i + 1, len(valid_loader))
        x, target = data[0].cuda(non_blocking=True), data[1]
        eng = net.engine(x.device)
        tgt = target[0, :, :, :155].to(x.device)
        tgt = torch.where(tgt == 4, torch.full_like(tgt, 3), tgt).to(torch.uint8)
        keeps = _po._keep_scales(net, x.device, 64).reshape(8, 8, 16)
        out = eng.predict_volume_tta(x.float(), keep_scales=keeps, target=tgt, want_probs=False, want_labels=True)
        soft = dice_from_counts(out["counts"].cpu().numpy())
        print(msg, soft)                                           # predict_cls.py:215
        wt_dices.append(soft[0]); tc_dices.append(soft[1]); et_dices.append(soft[2])
    print('WT Dice: %.4f' % np.mean(wt_dices))
    print('TC Dice: %.4f' % np.mean(tc_dices))
    print('ET Dice: %.4f' % np.mean(et_dices))
    return np.mean(wt_dices), np.mean(tc_dices), np.mean(et_dices)
