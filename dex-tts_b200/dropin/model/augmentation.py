"""``model.augmentation`` (DEX-TTS/model/augmentation.py; imported by the TRAINING data loader, DEX-TTS/src/dataset.py:11) is not part of
the CUDA inference package: fail with a message that says so instead of a bare ModuleNotFoundError."""
raise ImportError("model.augmentation (Augment) is training-only and is not provided by the dexb200 drop-in `model` package; "
                  "use the reference's own model/ directory for training")
