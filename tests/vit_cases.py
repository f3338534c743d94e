"""Shared ViT test cases: configs of the golden models and a loader for tests/golden/vit_*.npz."""
import os

import numpy as np
import torch

CASES = {
    "deepcnn_n": dict(embedding="deepresnet", embed_dim=64, num_heads=4, num_layers=6, activation="relu",
                      use_pos_encoding=False, use_regression_token=True),
    "linear_s_pos": dict(embedding="linear", embed_dim=32, num_heads=2, num_layers=3, activation="relu",
                         use_pos_encoding=True, use_regression_token=True),
    "cnn_s_mean": dict(embedding="cnn", embed_dim=32, num_heads=2, num_layers=3, activation="gelu",
                       use_pos_encoding=True, use_regression_token=False),
    "linear_s_feat_early": dict(embedding="linear", embed_dim=32, num_heads=2, num_layers=3, activation="relu",
                                use_pos_encoding=False, use_regression_token=True, use_global_features=True,
                                fusion_type="early"),
    "linear_s_feat_late": dict(embedding="linear", embed_dim=32, num_heads=2, num_layers=3, activation="relu",
                               use_pos_encoding=False, use_regression_token=True, use_global_features=True,
                               fusion_type="late"),
    # round 2 (oracle/make_golden_r2.py): the ImagesFeatures models (trainSettingsImagesFeatures.py:112-188) and F.leaky_relu
    "deepcnn_n_feat_early": dict(embedding="deepresnet", embed_dim=64, num_heads=4, num_layers=6, activation="relu",
                                 use_pos_encoding=False, use_regression_token=True, use_global_features=True, fusion_type="early"),
    "deepcnn_n_feat_late": dict(embedding="deepresnet", embed_dim=64, num_heads=4, num_layers=6, activation="relu",
                                use_pos_encoding=False, use_regression_token=True, use_global_features=True, fusion_type="late"),
    "linear_s_leaky": dict(embedding="linear", embed_dim=32, num_heads=2, num_layers=3, activation="leaky_relu",
                           use_pos_encoding=True, use_regression_token=True),
}
# ModularTransformer (helpers/models.py:366-593) goldens: tests/golden/vit_mod_*.npz (oracle/make_golden_modular.py)
MODULAR_CASES = {
    "mod_images_only_linear": dict(modular=True, mode="images_only", embedding="linear", embed_dim=32, num_heads=2, num_layers=2,
                                   hidden_dim=64, activation="relu", use_pos_encoding=False, use_regression_token=True),
    "mod_features_only_mlp": dict(modular=True, mode="features_only", embed_dim=32, num_heads=2, num_layers=2, hidden_dim=64,
                                  activation="relu", use_pos_encoding=True, use_regression_token=True, features_dim=6,
                                  feature_embedding_type="mlp"),
    "mod_both_add_linear": dict(modular=True, mode="both", embedding="linear", embed_dim=32, num_heads=2, num_layers=3,
                                hidden_dim=64, activation="relu", use_pos_encoding=True, use_regression_token=True, features_dim=5,
                                feature_embedding_type="linear", fusion_method="add"),
    "mod_both_concatproj_mlp_deep": dict(modular=True, mode="both", embedding="deepresnet", embed_dim=64, num_heads=4, num_layers=2,
                                         hidden_dim=128, activation="relu", use_pos_encoding=False, use_regression_token=False,
                                         features_dim=25, feature_embedding_type="mlp", fusion_method="concat_proj"),
    "mod_both_concatfeat_cnn": dict(modular=True, mode="both", embedding="cnn", embed_dim=32, num_heads=2, num_layers=2,
                                    hidden_dim=64, activation="gelu", use_pos_encoding=False, use_regression_token=True,
                                    features_dim=7, feature_embedding_type="linear", fusion_method="concat_features"),
    # per-frame outputs: no regression token, single_prediction=False (helpers/models.py:585-593); target / pred are [B, F, 1]
    "mod_perframe": dict(modular=True, mode="images_only", embedding="linear", embed_dim=32, num_heads=2, num_layers=2,
                         hidden_dim=64, activation="relu", use_pos_encoding=True, use_regression_token=False,
                         single_prediction=False),
}
PARAM_COUNTS = {"deepcnn_n": 506081, "linear_s_pos": 36865,    # SURVEY.md section 4 (reference notebooks)
                "deepcnn_n_feat_early": 511905, "deepcnn_n_feat_late": 520097}


def load_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "vit_%s.npz" % name))
    sd = {k[3:]: torch.tensor(z[k]) for k in z.files if k.startswith("sd/")}
    feats = torch.tensor(z["features"]) if "features" in z.files else None
    x = torch.tensor(z["x"]) if "x" in z.files else None       # 'features_only' ModularTransformer cases have no images
    return z, sd, x, torch.tensor(z["target"]), feats


