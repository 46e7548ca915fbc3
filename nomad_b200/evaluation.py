"""Evaluation-harness entry points of the reference (``src/training/train_triplet.py:203-474``) on the B200 path.

The reference's harness re-uses the scoring hot path -- ``get_embeddings_csv`` (``:203-225``), ``cdist`` + row mean
(``:267-268, 322-323, 374-375``), the ``cdist`` diagonal of full-reference mode (``:438-439``) and condition-grouped
means (``:274, 381, 445``) -- around plots and prints.  This module binds exactly those numeric steps to the C ABI
(batched windowed embedding, ``nomad_b200_cdist_mean``, ``nomad_b200_paired_dist``) under the reference's method
names and config keys, and returns the tables the reference prints or plots (the seaborn figures themselves are out
of scope).  Nothing here trains; see ``triplet.py`` for the fine-tuning step.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import pandas as pd
import torch


def order_three(x, a, b, c, d):
    """``Training.order_three`` (``train_triplet.py:227-228``)."""
    return a * x + b * x ** 2 + c * x ** 3 + d


class Evaluation:
    """``Training``'s evaluation half (``train_triplet.py:203-489``).  ``nomad`` is a :class:`nomad_b200.nomad.Nomad`;
    ``config`` uses the reference's keys (``src/config/train_triplet.yaml``): ``non_match_dir``, ``test_db_file``,
    ``test_root_wav``, ``db``, ``conds``, ``test_db_file_fr``, ``test_mono_data``, ``test_mono_wav``."""

    def __init__(self, nomad, config: Dict):
        self.nomad = nomad
        self.config = dict(config)

    # ---- train_triplet.py:203-225
    def get_embeddings_csv(self, model, file_names, root=False):
        """Same frame as the reference: the file-name column followed by columns 0..255 (one batched pass)."""
        emb = self.nomad.embed_files(np.array(file_names), root).cpu().numpy()
        return pd.concat([file_names.reset_index(), pd.DataFrame(emb)], axis=1).drop('index', axis=1)

    # ---- train_triplet.py:476-484
    def get_nmr_embeddings(self):
        ref_files = pd.DataFrame(os.listdir(self.config['non_match_dir']))
        ref_files.columns = ['reference']
        ref_files['reference'] = [os.path.join(self.config['non_match_dir'], x) for x in ref_files['reference']]
        return self.get_embeddings_csv(self.nomad.model, ref_files['reference'])

    def _nmr_distance(self, test_embeddings: pd.DataFrame, ref_embeddings: pd.DataFrame) -> np.ndarray:
        """``np.mean(cdist(test, ref), axis=1)`` (``:267-268``) on the GPU; float64 like numpy's."""
        _, avg = self.nomad.pairwise(test_embeddings, ref_embeddings)
        return avg

    @staticmethod
    def _correlations(df_dist: pd.DataFrame, target: str) -> Dict[str, float]:
        """SRCC / PCC with and without the third-order mapping (``:279-306``)."""
        from scipy.optimize import curve_fit
        from scipy.stats import pearsonr, spearmanr
        out = {}
        out['SRCC'] = float(spearmanr(df_dist['Distance'], df_dist[target])[0])
        out['PCC'] = float(pearsonr(df_dist['Distance'], df_dist[target])[0])
        try:
            popt3, _ = curve_fit(order_three, df_dist['Distance'].values, df_dist[target].values)
            df_dist['Distance_map'] = df_dist['Distance'].apply(lambda x: order_three(x, *popt3))
            out['SRCC_map'] = float(spearmanr(df_dist['Distance_map'], df_dist[target])[0])
            out['PCC_map'] = float(pearsonr(df_dist['Distance_map'], df_dist[target])[0])
        except (RuntimeError, TypeError):  # fewer conditions than polynomial coefficients
            pass
        return out

    # ---- train_triplet.py:231-306
    def eval_audio_quality(self, test_data: Optional[pd.DataFrame] = None):
        """Per database: NMR distance of every test file, averaged per condition, against MOS.
        -> {db_name: (df_dist indexed by condition [Distance, mos, (Distance_map)], {'SRCC', 'PCC', ...})}"""
        if test_data is None:
            test_data = pd.read_csv(self.config['test_db_file'])
        if self.config.get('db') is not None:
            test_data = test_data[test_data['db'].isin(self.config['db'])]
        if self.config.get('conds') is not None:
            test_data = test_data[test_data['condition'].str.contains('|'.join(self.config['conds']))]
        ref_embeddings = self.get_nmr_embeddings().set_index('reference')
        out = {}
        for db_name, db in test_data.groupby('db'):
            df_emb = self.get_embeddings_csv(self.nomad.model, db['filepath_deg'], root=self.config.get('test_root_wav', False))
            test_embeddings = df_emb.set_index('filepath_deg')
            test_names = df_emb.merge(db, on='filepath_deg')[['filepath_deg', 'condition', 'mos']]
            avg_dist_nmr = self._nmr_distance(test_embeddings, ref_embeddings)
            df_dist = pd.DataFrame({'filepath_deg': test_embeddings.index, 'Distance': avg_dist_nmr})
            df_dist = df_dist.merge(test_names, on='filepath_deg').set_index('filepath_deg')
            df_dist = df_dist.groupby('condition').mean()
            out[db_name] = (df_dist, self._correlations(df_dist, 'mos'))
        return out

    # ---- train_triplet.py:308-345
    def eval_degr_level(self, anchors: pd.Series, root=False):
        """Validation anchors sorted by NMR distance, with the reference's condition label parsed from the file name."""
        df_emb = self.get_embeddings_csv(self.nomad.model, anchors, root=root)
        ref_embeddings = self.get_nmr_embeddings()
        avg = self._nmr_distance(df_emb.iloc[:, 1:], ref_embeddings.iloc[:, 1:])
        df_dist = pd.DataFrame({'Anchor': df_emb.iloc[:, 0], 'Distance': avg})
        df_dist.sort_values(by='Distance', inplace=True)
        df_dist['condition'] = [x.split('_')[1] + ' ' + x.split('_')[2].split('.')[0] for x in df_dist['Anchor']]
        order = df_dist.groupby('condition')['Distance'].mean().sort_values().index
        return df_dist, list(order)

    # ---- train_triplet.py:347-417
    def eval_degradation_intensity(self, test_data: Optional[pd.DataFrame] = None):
        """Per degradation: condition-averaged NMR distance and its SRCC with the intensity level."""
        from scipy.stats import spearmanr
        if test_data is None:
            test_data = pd.read_csv(self.config['test_mono_data'])
        ref_embeddings = self.get_nmr_embeddings().set_index('reference')
        out = {}
        for deg_name, deg_data in test_data.groupby('Degradation'):
            df_emb = self.get_embeddings_csv(self.nomad.model, deg_data['filepath_deg'], root=self.config.get('test_mono_wav', False))
            test_embeddings = df_emb.set_index('filepath_deg')
            test_names = df_emb.merge(deg_data, on='filepath_deg')[['filepath_deg', 'Condition']]
            avg = self._nmr_distance(test_embeddings, ref_embeddings)
            df_dist = pd.DataFrame({'filepath_deg': test_embeddings.index, 'Distance': avg}).merge(test_names, on='filepath_deg')
            df_dist.set_index('filepath_deg', inplace=True)
            df_dist = df_dist.groupby('Condition').mean().reset_index()
            df_dist.sort_values(by='Distance', inplace=True)
            out[deg_name] = (df_dist, float(spearmanr(df_dist['Distance'], df_dist['Condition'])[0]))
        return out

    # ---- train_triplet.py:419-474
    def eval_full_reference(self, test_data: Optional[pd.DataFrame] = None):
        """Distance of every test file to its OWN reference: the reference takes ``np.diag(cdist(test, ref))``
        (``:438-439``); here the diagonal is computed directly (``nomad_b200_paired_dist``), never the matrix."""
        if test_data is None:
            test_data = pd.read_csv(self.config['test_db_file_fr'])
        root = self.config.get('test_root_wav', False)
        out = {}
        for db_name, db in test_data.groupby('db'):
            df_emb_ref = self.get_embeddings_csv(self.nomad.model, db['filepath_ref'], root=root).set_index('filepath_ref')
            df_emb_test = self.get_embeddings_csv(self.nomad.model, db['filepath_deg'], root=root).set_index('filepath_deg')
            test_names = df_emb_test.merge(db, on='filepath_deg')[['filepath_deg', 'condition', 'mos']]
            a = torch.from_numpy(np.ascontiguousarray(df_emb_test.to_numpy(dtype=np.float32)))
            b = torch.from_numpy(np.ascontiguousarray(df_emb_ref.to_numpy(dtype=np.float32)))
            fr_distance = self.nomad.engine.paired_dist(a, b).cpu().numpy()
            df_dist = pd.DataFrame({'filepath_deg': df_emb_test.index, 'Distance': fr_distance})
            df_dist = df_dist.merge(test_names, on='filepath_deg')
            df_dist = df_dist.groupby('condition').mean(numeric_only=True)
            out[db_name] = (df_dist, self._correlations(df_dist, 'mos'))
        return out

    # ---- train_triplet.py:486-489
    @staticmethod
    def euclidean_dist(emb_a, emb_b):
        return np.sqrt(np.dot(emb_a - emb_b, (emb_a - emb_b).T))
