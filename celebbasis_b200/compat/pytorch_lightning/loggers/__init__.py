"""Logger stand-ins (main_id_embed.py:640-652, ImageLogger's `pl.loggers.TestTubeLogger` key at :353)."""
class _NullExperiment:
    def __getattr__(self, name):
        return lambda *a, **k: None


class LightningLoggerBase:
    def __init__(self, save_dir=None, name="default", version=None, **kwargs):
        self.save_dir, self.name, self.version = save_dir, name, version
        self.experiment = _NullExperiment()

    def log_metrics(self, metrics, step=None):
        pass

    def log_hyperparams(self, params):
        pass

    def save(self):
        pass

    def finalize(self, status):
        pass


class TestTubeLogger(LightningLoggerBase):
    pass


class CSVLogger(LightningLoggerBase):
    pass


class WandbLogger(LightningLoggerBase):
    pass
