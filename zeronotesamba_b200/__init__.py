"""zeronotesamba_b200 -- B200-native hot path of deezer/zeroNoteSamba (VQT front-end + two-branch
Down_CNN pretext training step) behind the reference's Python API.  See DESIGN.md.

    import zeronotesamba_b200.processing.input_rep as IR      # IR.generate_XQT(y, 16000, "vqt")
    from zeronotesamba_b200.models.models import Down_CNN, Pretext_CNN
    from zeronotesamba_b200.models.loss_functions import NTXent
    from zeronotesamba_b200.pretext import train_epoch, val_epoch, PretextTrainer, FusedAdam
"""
__version__ = "0.1.0"
