"""Plot interceptors (reference src/intercepts/plot_intercepts.py:24-162).  They are consumers
of `encoding_prediction` / `generate_cells` only and are not wired into `python3 src`
(src/__main__.py:53-66).  matplotlib / umap-learn are not installed in this image, so the
heavy imports are deferred: constructing the object works, saving a figure raises a clear
ImportError when matplotlib is missing."""
import os


def _pyplot():
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
        return plt
    except ImportError as exc:  # pragma: no cover - depends on the image
        raise ImportError('PlotIntercepts needs matplotlib, which is not installed') from exc


class PlotIntercepts:
    def __init__(self, log_dir, trainer=None):
        self.log_dir = log_dir
        self.trainer = trainer
        self.plot_dir = os.path.join(log_dir, 'plots')

    def _encodings(self):
        return self.trainer.network.encoding_prediction(self.trainer.data)

    def save_encoding_plot(self, prefix='enc'):
        """Scatter of the first two (or three) encoding dimensions of all cells, one PNG per
        iteration under <log_dir>/plots/."""
        def intercept(it, _):
            plt = _pyplot()
            os.makedirs(self.plot_dir, exist_ok=True)
            enc = self._encodings()
            fig = plt.figure(figsize=(8, 8))
            if enc.shape[1] >= 3:
                ax = fig.add_subplot(projection='3d')
                ax.scatter(enc[:, 0], enc[:, 1], enc[:, 2], s=2)
            else:
                ax = fig.add_subplot()
                ax.scatter(enc[:, 0], enc[:, 1 % enc.shape[1]], s=2)
            ax.set_title(f'iteration {it}')
            fig.savefig(os.path.join(self.plot_dir, f'{prefix}_{it:06}.png'))
            plt.close(fig)

        return intercept

    def save_generated_cells_plot(self, samples=16, prefix='gen'):
        def intercept(it, _):
            plt = _pyplot()
            os.makedirs(self.plot_dir, exist_ok=True)
            net = self.trainer.network
            cells = net.generate_cells(net.random_encoding_vector(samples))
            fig, ax = plt.subplots(figsize=(12, 4))
            ax.imshow(cells, aspect='auto', interpolation='nearest')
            ax.set_title(f'generated cells, iteration {it}')
            fig.savefig(os.path.join(self.plot_dir, f'{prefix}_{it:06}.png'))
            plt.close(fig)

        return intercept
