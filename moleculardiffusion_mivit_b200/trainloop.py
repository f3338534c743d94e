"""The experiment loop around the hot path: dataset refresh per cycle, training of every
model of the experiment, per-cycle validation and the results / weights file -- the flat script
Experiments/PSFNoise/trainModelsPSFNoise.py:14-22,113-251 (same skeleton in the Framerate / Embeddings /
ImagesFeatures trainers) as a reusable object.

What is kept from the reference, line by line:
  * batch size 1, doubled every `adaptive_batch_size` cycles (:117-119; -1 -> fixed 16, :40);
  * per cycle and per D group `[mean, var]` of `TrainingDs_list`: N trajectories (N // 2 for the 10.2 group, :128) of T
    sub-steps, divided by `traj_div_factor` (:154), rendered by the experiment's renderer (:155); labels D / D_max (:164);
  * every model sees every (shuffled) batch of the cycle: zero_grad, forward, MSELoss, backward, AdamW step (:177-193),
    then ONE StepLR(5, 0.9) step per cycle (:196);
  * validation in eval mode on the fixed sets: MSE(pred * D_max, D) per set and their mean (:206-238) appended to
    `validation_losses[name]["val_<D>" | "val_avg"]`;
  * `save_results` writes {"validation_losses", "all_labels", "model_weights": {name: state_dict}} with torch.save for the
    last five cycles (suffix = cycles remaining, :241-242) and at the end (:251) -- files load into the reference's models.
ImagesFeatures variant (Experiments/ImagesFeatures/trainModelsImagesFeatures.py:155-203): when `render_fn` returns
(videos, features) -- e.g. a wrapper of create_video_and_feature_pairs -- the features travel with the videos through the
shuffle and every model is called the way that trainer calls it: "im_resnet" -> model(images), "ft_mlp" -> model(features),
names without "ft" -> model(images, None), the rest -> model(images, features) (:187-199); validation sets are then
(images, features, D) triples (make_prediction_tuple, :226).
What changes: trajectories come from the device Brownian generator (`mivit_brownian`; the reference calls the third-party
andi_datasets `models_phenom().single_state(N, L=0, T, Ds=[mean, var], alphas=1)`, statistically equivalent),
the optimiser step is the fused `MiViTTrainer.train_step`, and everything stays on the GPU between rendering and training."""
import inspect

import numpy as np
import torch

from . import helpersGeneration as _gen
from .models import _CudaViT
from .training import MiViTTrainer

__all__ = ["save_results", "single_state", "ExperimentLoop"]


def save_results(validation_losses, all_gen_labels, models, path_addition="", prefix="training_results_PSFNoise"):
    """trainModelsPSFNoise.py:14-22."""
    save_path = prefix + path_addition + ".pth"
    results = {"validation_losses": validation_losses, "all_labels": all_gen_labels,
               "model_weights": {name: model.state_dict() for name, model in models.items()}}
    torch.save(results, save_path)
    print(f"\nTraining results saved to {save_path}")
    return save_path


def single_state(N, T, Ds, seed=None, seq_offset=0):
    """Stand-in for `models_phenom().single_state(N, L=0, T=T, Ds=[mean, var], alphas=1)` + the transposes of
    trainModelsPSFNoise.py:128-135: returns (trajs (N,T,2) float64 numpy in pixels, D (N,) float32)."""
    traj, D = _gen.brownian_motion(N, T, 1, [Ds[0]], 1.0, seed=seed, seq_offset=seq_offset, D_var=Ds[1], return_D=True)
    return traj, D


class _TorchTrainer:
    """Same interface as MiViTTrainer for models OUTSIDE this package (e.g. the ImagesFeatures experiment's feature-only MLP
    `ft_mlp`, an nn.Sequential): same criterion, optimiser and schedule (trainSettingsPSFNoise.py:119-120) on stock PyTorch CUDA.
    The reference's own models -- the ViTs and the CNN baselines MultiImageResNet / MultiImageFeatureResNet -- have CUDA-library
    trainers (training.MiViTTrainer, baselines.CnnTrainer)."""

    def __init__(self, model, lr=1e-4, step_size=5, gamma=0.9):
        self.model = model.cuda()
        self.opt = torch.optim.AdamW(self.model.parameters(), lr=lr)
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, step_size=step_size, gamma=gamma)
        self.criterion = torch.nn.MSELoss()

    def train_step(self, x, target, features=None):
        self.model.train()
        self.opt.zero_grad()
        if x is None:
            out = self.model(features)
        else:
            out = self.model(x) if features is None else self.model(x, features)
        loss = self.criterion(out, target)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def scheduler_step(self):
        self.sched.step()


class ExperimentLoop:
    def __init__(self, models, render_fn, make_prediction, val_sets, *, T, N=64, traj_div_factor=100, D_max_normalization=10,
                 TrainingDs_list=((1, 1), (3, 1), (5, 1), (7, 1), (9, 1), (10.2, 1)), adaptive_batch_size=20, shuffle=True,
                 lr=1e-4, step_size=5, gamma=0.9, seed=0, results_prefix="training_results_PSFNoise"):
        """models: {name: GeneralTransformer / ModularTransformer / CNN baseline of this package (fused CUDA trainers) or any other
        nn.Module (stock PyTorch trainer)}; render_fn(trajs numpy (n,T,2)[, seq_offset=global id of trajs[0]][, seed=...]) -> images
        (numpy or CUDA tensor, sequence axis first) or (images, features); make_prediction(model, name, images[, features]) ->
        predictions (the settings files' function, e.g. images[:, psf, noise] for PSFNoise); val_sets: [(images, D_value), ...] or
        [(images, features, D_value), ...] rendered once (load_validation_data).
        Noise streams: the renderer keys them by (seed, global sequence id); render_fn is handed the global id of its first
        trajectory (and the loop's seed) whenever its signature accepts `seq_offset` (`seed`), so every D group and every cycle
        draws fresh noise like the reference does -- a render_fn that takes neither must vary its seed itself."""
        self.models, self.render_fn, self.make_prediction = models, render_fn, make_prediction
        self.val_sets = []
        for vs in val_sets:
            feats = torch.as_tensor(vs[1]).float().cuda() if len(vs) == 3 else None
            self.val_sets.append((torch.as_tensor(vs[0]).float().cuda(), feats, float(vs[-1])))
        try:
            prm = inspect.signature(render_fn).parameters
            anykw = any(p.kind == p.VAR_KEYWORD for p in prm.values())
            self._render_kw = {k for k in ("seq_offset", "seed") if anykw or k in prm}
        except (TypeError, ValueError):
            self._render_kw = set()
        self.T, self.N, self.div, self.Dmax = int(T), int(N), float(traj_div_factor), float(D_max_normalization)
        self.groups = [list(g) for g in TrainingDs_list]
        self.adaptive, self.shuffle, self.seed = int(adaptive_batch_size), bool(shuffle), int(seed)
        self.batch_size = 1 if self.adaptive != -1 else 16
        self.trainers = {n: self._make_trainer(m, lr, step_size, gamma) for n, m in models.items()}
        self.validation_losses = {n: dict({f"val_{d}": [] for _, _, d in self.val_sets}, val_avg=[]) for n in models}
        self.all_gen_labels = np.array([])
        self.prefix = results_prefix
        self._generated = 0
        self._gen = torch.Generator(device="cuda").manual_seed(self.seed)

    @staticmethod
    def _make_trainer(m, lr, step_size, gamma):
        from .baselines import CnnTrainer, _CudaResNet
        if isinstance(m, _CudaViT):
            return MiViTTrainer(m, lr=lr, step_size=step_size, gamma=gamma)
        if isinstance(m, _CudaResNet):        # the experiments' CNN baseline (helpers/models.py:686 MultiImageResNet & co.)
            return CnnTrainer(m, lr=lr, step_size=step_size, gamma=gamma)
        return _TorchTrainer(m, lr=lr, step_size=step_size, gamma=gamma)

    # -- one dataset refresh (:121-173) ---------------------------------------------------------------------------
    def generate(self):
        videos, feats, labels = [], [], []
        for g in self.groups:
            n = self.N if g[0] != 10.2 else self.N // 2
            first_id = self._generated
            trajs, D = single_state(n, self.T, g, seed=self.seed, seq_offset=first_id)
            self._generated += n
            self.all_gen_labels = np.append(self.all_gen_labels, D)
            kw = {}
            if "seq_offset" in self._render_kw:
                kw["seq_offset"] = first_id
            if "seed" in self._render_kw:
                kw["seed"] = self.seed
            out = self.render_fn(trajs / self.div, **kw)
            if isinstance(out, (tuple, list)):                 # ImagesFeatures: (videos, features[, ...])
                feats.append(torch.as_tensor(out[1]).float().cuda())
                out = out[0]
            videos.append(torch.as_tensor(out).float().cuda())
            labels.append(torch.from_numpy(D / self.Dmax))
        return (torch.cat(videos, 0), torch.cat(feats, 0) if feats else None,
                torch.cat(labels, 0).float().unsqueeze(-1).cuda())

    def train_cycle(self, cycle):
        if self.adaptive != -1 and cycle != 0 and cycle % self.adaptive == 0:
            self.batch_size *= 2
        videos, feats, labels = self.generate()
        n = videos.shape[0]
        order = torch.randperm(n, device="cuda", generator=self._gen) if self.shuffle else torch.arange(n, device="cuda")
        last = {}
        for name, model in self.models.items():
            model.train()
            tr = self.trainers[name]
            for i in range(0, n, self.batch_size):
                idx = order[i:i + self.batch_size]
                x, f = self._select(model, name, videos[idx], feats[idx] if feats is not None else None)
                last[name] = tr.train_step(x, labels[idx], f).clone()
            tr.scheduler_step()
        return {k: float(v.item()) for k, v in last.items()}

    def _select(self, model, name, images, features=None):
        """What the model is called with.  Image-only experiments: the settings files' make_prediction both slices the image
        stack and calls the model; the fused trainer needs the slice only, so the model call is intercepted.  With features
        (ImagesFeatures) the trainer's own rule applies (trainModelsImagesFeatures.py:187-199)."""
        if features is not None:
            if name == "im_resnet":
                return images.contiguous(), None
            if name == "ft_mlp":
                return None, features.contiguous()
            return images.contiguous(), (features.contiguous() if "ft" in name else None)
        box = {}

        class _Probe:
            def __call__(self_inner, inp, *a):
                box["x"], box["extra"] = inp, a
                return inp

            def eval(self_inner):
                return self_inner

        self.make_prediction(_Probe(), name, images)
        return box["x"].contiguous(), None

    def validate(self):
        for name, model in self.models.items():
            model.eval()
            with torch.no_grad():
                losses = []
                for vid, vfeat, d in self.val_sets:
                    label = torch.full((vid.shape[0], 1), d, device="cuda")
                    if vfeat is not None:
                        pred = self.make_prediction(model, name, vid, vfeat) * self.Dmax
                    else:
                        pred = self.make_prediction(model, name, vid) * self.Dmax
                    loss = torch.nn.functional.mse_loss(pred, label).item()
                    self.validation_losses[name][f"val_{d}"].append(loss)
                    losses.append(loss)
                self.validation_losses[name]["val_avg"].append(float(np.mean(losses)))

    def run(self, num_cycles, save=True):
        for cycle in range(num_cycles):
            self.train_cycle(cycle)
            self.validate()
            if save and num_cycles - cycle - 1 < 5:
                save_results(self.validation_losses, self.all_gen_labels, self.models, str(num_cycles - cycle), self.prefix)
        if save:
            save_results(self.validation_losses, self.all_gen_labels, self.models, "", self.prefix)
        return self.validation_losses
