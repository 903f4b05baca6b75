"""Multi-modal Multi-task Transformer for Captioning -- drop-in for the reference's
model/MMT4Caption.py (same constructor, ``mode`` / ``forward`` / ``greedy_decode`` API, config-JSON
surface and state_dict keys), running the caption task on the B200 kernels of libvct_b200.so.

Hot path: ``caption_forward`` (tokenise -> MultiModalEncoder -> CapDecoder + SCE loss,
reference :114-121) and ``greedy_decode`` (:146-184, here K/V-cached and with device-side argmax/end
flags).  "match"/"cross" need the frozen CLIP/BERT text tower and are outside the hot path; the methods
exist and work when that tower is available."""
import os
import weakref
from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn
from torch import Tensor

from .CapDecoder import CapDecoder
from .CapPreprocessor import CapPreprocessor
from .TextEncoder import TextEncoder
from .MMEncoder import MultiModalEncoder, SimpleSepEncoder, HMMEncoder
from .Matching import Matching
from ._engine import build_engine, param_device


class MMT4Caption(nn.Module):
    def __init__(self, model_config: dict, device=torch.device('cuda')):
        """:param model_config: the "model" block of the JSON config (configs/*.json)
        :param device: torch.device"""
        super().__init__()
        self.device = device
        self.model_config = model_config
        self.loss_beta = model_config['loss_beta']
        self.f_type = None

        # construction order = the reference's (:29-91), so a given torch seed yields identical weights
        self.cap_preprocessor = CapPreprocessor(model_config['tokenizer'], device=device)
        self.text_encoder = TextEncoder(model_config['text_enc_type'], device=device)
        self.cap_decoder = CapDecoder(
            num_layers=model_config['caption_decoder']['layer'],
            embed_dim=model_config['embed_dim'],
            nhead=model_config['caption_decoder']['nhead'],
            dim_feedforward=model_config['caption_decoder']['feedforward'],
            dropout=model_config['dropout'],
            vocab_size=self.cap_preprocessor.tokenizer.vocab_size,
            pad_id=self.cap_preprocessor.pad_id,
            sce_loss_alpha=model_config['caption_decoder']['sce_loss_alpha'],
            custom_decoder_type=model_config['caption_decoder'].get('layer_type', None),
            activation=model_config['activation'],
            device=device
        )
        vid_enc_type = model_config['video_encoder'].get('type', 'mme')
        enc_cls = {"simple": SimpleSepEncoder, "hmme": HMMEncoder}.get(vid_enc_type)
        if enc_cls is not None:
            self.video_encoder = enc_cls()       # raises: outside the hot path
        else:
            mme = model_config['video_encoder']['mme']
            self.video_encoder = MultiModalEncoder(
                d_feats=model_config['modal_shape'],
                d_model=model_config['embed_dim'],
                nhead=model_config['video_encoder']['nhead'],
                dim_feedforward=model_config['video_encoder']['feedforward'],
                num_encoder_layers=model_config['video_encoder']['layer'],
                dropout=model_config['dropout'],
                activation=model_config['activation'],
                global_type=mme['aggregation'],
                modal_different=mme.get('modal_different', True),
                temporal_type=mme.get('temporal', 'encoding'),
                do_norm=mme.get('do_norm', False),
                device=device
            )
        if model_config.get('matching', None) is not None:
            self.matching = Matching((model_config['embed_dim'], self.text_encoder.dim),
                                     enable_tem=model_config['matching']['enable_tem'],
                                     loss=model_config['matching']['matching_loss'],
                                     loss_tem=model_config['matching'].get("temperature", None),
                                     device=device)
        self._vct_engine = None
        ref = weakref.ref(self)
        object.__setattr__(self.cap_decoder, "_vct_owner", ref)
        object.__setattr__(self.video_encoder, "_vct_owner", ref)

    # ---- engine ---------------------------------------------------------------------------------
    def _engine(self):
        eng = self._vct_engine
        if eng is None or not eng.arena.is_current():
            dev = param_device(self.cap_decoder)
            eng = build_engine(self.video_encoder, self.cap_decoder, dev, precision=self.__dict__.get("vct_precision"), gemm_impl=self.__dict__.get("vct_gemm"))
            object.__setattr__(self, "_vct_engine", eng)
            object.__setattr__(self, "device", dev)
            self.cap_preprocessor.device = dev
        return eng

    def reset_engine(self):
        object.__setattr__(self, "_vct_engine", None)

    # ---- forward dispatch (reference :96-112) -----------------------------------------------------
    def forward(self, video_feats: List[Tensor], video_masks: List[Tensor], captions: Union[List[str], Tensor]):
        """video_feats: [Tensor[B,T,E]] per modality; video_masks: [Bool[B,T]] (True = padded);
        captions: raw strings (reference API) or an int64 id tensor [B, L] with [CLS] ... [SEP] [PAD]*
        (lets synthetic benchmarks bypass the host tokenizer)."""
        if self.f_type == "caption":
            return self.caption_forward(video_feats, video_masks, captions)
        elif self.f_type == "match":
            return self.match_forward(video_feats, video_masks, captions)
        elif self.f_type == "cross":
            return self.cross_forward(video_feats, video_masks, captions)
        else:
            raise ValueError

    def _tokenise(self, captions):
        if isinstance(captions, Tensor):
            ids = captions.to(self._engine().device, non_blocking=True)
            return ids, ids == self.cap_preprocessor.pad_id
        return self.cap_preprocessor(list(captions))

    def caption_forward(self, video_feats, video_masks, captions):
        """loss of the captioning task (reference :114-121; returns the loss only)."""
        eng = self._engine()
        text_ts, text_mask_ts = self._tokenise(captions)
        self.video_encoder._vct_S_hint = text_ts.shape[1] - 1
        # one fresh dropout step per training forward, shared by the encoder and the decoder (distinct call sites)
        eng._pinned_step = None
        if self.training:
            eng._pinned_step = eng.next_rng_step()
        try:
            memory, _, _ = self.video_encoder(video_feats, video_masks)
            # the logits are discarded here (reference :120-121), so the decoder does not copy them out
            _, loss = self.cap_decoder(memory, text_ts, text_mask_ts, _want_logits=False)
        finally:
            eng._pinned_step = None
        return loss

    def match_forward(self, video_feats, video_masks, captions):
        text_feat = self.text_encoder(captions)
        _, _, agg_feat = self.video_encoder(video_feats, video_masks)
        return self.matching(text_feat, agg_feat)

    def cross_forward(self, video_feats, video_masks, captions):
        text_ts, text_mask_ts = self._tokenise(captions)
        text_feat = self.text_encoder(captions)
        self.video_encoder._vct_S_hint = text_ts.shape[1] - 1
        memory, memory_masks, agg_feat = self.video_encoder(video_feats, video_masks)
        logits, cap_loss = self.cap_decoder(memory, text_ts, text_mask_ts)
        match_loss = self.matching(text_feat, agg_feat)
        loss = self.loss_beta * cap_loss + (1 - self.loss_beta) * match_loss
        return loss, cap_loss, match_loss

    # ---- decoding (reference :146-184) -------------------------------------------------------------
    def greedy_decode_ids(self, video_feat: List[Tensor], video_masks: Optional[List[Tensor]] = None,
                          max_len: int = 30, sync_every: int = 1) -> Tensor:
        """Token ids [B, n] incl. the leading [CLS] (device tensor)."""
        eng = self._engine()
        if self.cap_decoder._uses_patched_layers():
            return self._greedy_patched(video_feat, video_masks, max_len)
        feats = video_feat[0].to(eng.device)
        mask = video_masks[0].to(eng.device) if video_masks is not None else None
        return eng.greedy_decode(feats, mask, max_len, self.cap_preprocessor.start_id, self.cap_preprocessor.end_id,
                                 sync_every=sync_every, eval_fastpath=not self.training)

    def _greedy_patched(self, video_feat, video_masks, max_len):
        """predict_video.py:43-79,126-130 rebinds every decoder layer's ``forward`` to capture
        ``self.mha`` (head-averaged cross-attention maps [B, L, M]).  The incremental decoder produces the
        same maps natively and stores them where the patched forward would."""
        eng = self._engine()
        feats = video_feat[0].to(eng.device)
        mask = video_masks[0].to(eng.device) if video_masks is not None else None
        ys, probs = eng.greedy_decode(feats, mask, max_len, self.cap_preprocessor.start_id,
                                      self.cap_preprocessor.end_id, want_probs=True, eval_fastpath=not self.training)
        for layer, pr in zip(self.cap_decoder.decoder.layers, probs):
            layer.mha = pr
        return ys

    def greedy_decode(self, video_feat: List[Tensor], video_masks: Optional[List[Tensor]] = None,
                      max_len: int = 30) -> List[str]:
        ys = self.greedy_decode_ids(video_feat, video_masks, max_len)
        end_id = self.cap_preprocessor.end_id
        result = []
        for idx_cap in ys.tolist():            # single device->host transfer for the whole batch
            end_count = -1
            for i, idx in enumerate(idx_cap):
                if idx == end_id:
                    end_count = i
                    break
            idx_cap = idx_cap[1:end_count]      # no [SEP]: end_count = -1 drops the last token (Q11)
            token_cap = self.cap_preprocessor.tokenizer.convert_ids_to_tokens(idx_cap)
            result.append(self.cap_preprocessor.tokenizer.convert_tokens_to_string(token_cap))
        return result

    def beam_decode(self):
        pass

    def mode(self, forward_type="caption") -> None:
        """"caption", "match" or "cross" (reference :189-211): sets the task and which heads train."""
        flags = {"caption": (True, False), "match": (False, True), "cross": (True, True)}
        if forward_type not in flags:
            raise ValueError
        self.f_type = forward_type
        dec, mat = flags[forward_type]
        for param in self.cap_decoder.parameters():
            param.requires_grad = dec
        for param in self.matching.parameters():
            param.requires_grad = mat

    # ---- weight importers (reference :213-283) ------------------------------------------------------
    def load_embedding_from_bert(self):
        from transformers import BertModel
        bert = BertModel.from_pretrained("bert-base-uncased")
        with torch.no_grad():
            self.cap_decoder.tgt_to_emb.weight.copy_(bert.embeddings.word_embeddings.weight)
            n = bert.embeddings.position_embeddings.weight.shape[0]
            self.cap_decoder.positional_encoding.pos_embedding[:n].copy_(bert.embeddings.position_embeddings.weight)

    def load_cap_decoder_from_univl(self, path):
        """Map a UniVL checkpoint's decoder onto cap_decoder (key table of reference :213-283)."""
        univl: Dict[str, Tensor] = torch.load(path, map_location="cpu")
        out: Dict[str, Tensor] = {}
        n_layers = len(self.cap_decoder.decoder.layers)
        for l in range(n_layers):
            src, dst = f'decoder.decoder.layer.{l}.', f'decoder.layers.{l}.'
            for wb in ('weight', 'bias'):
                for ours, theirs in (('self_attn', 'slf_attn'), ('multihead_attn', 'enc_attn')):
                    out[f'{dst}{ours}.in_proj_{wb}'] = torch.cat(
                        [univl[f'{src}{theirs}.att.{part}.{wb}'] for part in ('query', 'key', 'value')], dim=0)
                    out[f'{dst}{ours}.out_proj.{wb}'] = univl[f'{src}{theirs}.output.dense.{wb}']
                out[f'{dst}norm1.{wb}'] = univl[f'{src}slf_attn.output.LayerNorm.{wb}']
                out[f'{dst}norm2.{wb}'] = univl[f'{src}enc_attn.output.LayerNorm.{wb}']
                out[f'{dst}linear1.{wb}'] = univl[f'{src}intermediate.dense.{wb}']
                out[f'{dst}linear2.{wb}'] = univl[f'{src}output.dense.{wb}']
                out[f'{dst}norm3.{wb}'] = univl[f'{src}output.LayerNorm.{wb}']
        for wb in ('weight', 'bias'):
            out[f'decoder.norm.{wb}'] = univl[f'decoder.embeddings.LayerNorm.{wb}']
        out['generator.weight'] = univl['decoder.classifier.cls.predictions.decoder.weight']
        out['generator.bias'] = univl['decoder.classifier.cls.predictions.bias']
        out['tgt_to_emb.weight'] = univl['decoder.embeddings.word_embeddings.weight']
        out['positional_encoding.pos_embedding'] = univl['decoder.embeddings.position_embeddings.weight']
        self.cap_decoder.load_state_dict(out)
