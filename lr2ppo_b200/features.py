"""On-the-fly feature extraction: the ViT-B/16 and RoBERTa-base towers produce the
`clean_feat.h5`-shaped tensors the stage scripts read from disk (format: finetune/pointwise.py:136-143 —
`text_emb` float [n_tags, 196, 768], `img_emb` float [1, n_img, 768]) directly on the GPU.

The reference never runs the towers itself (the h5 file is produced offline); the conventions here follow
SURVEY.md §8(d) config 2: the real tokens of a tag are padded to the fusion model's 196 text positions with
`seg = 0` on the padding (masked keys, tencentpretrain/encoders/transformer_encoder.py:62-68), an image is
represented by the CLS position of the ViT output, and `pad_images` reproduces the dataset's cyclic padding of the
keyframe vectors to `max_imgs` (finetune/pointwise.py:139-154).
"""
import torch


class FeatureExtractor:
    def __init__(self, vit_model, text_model, seq_length=196, max_imgs=16, pad_id=1):
        """vit_model / text_model: lr2ppo_b200.tower.build_model(...) instances (embedding + encoder)."""
        self.vit, self.text = vit_model, text_model
        self.seq_length, self.max_imgs, self.pad_id = seq_length, max_imgs, pad_id

    @torch.no_grad()
    def text_features(self, tokens, lengths=None):
        """tokens int64 [n_tags, L] (L <= seq_length), lengths [n_tags] real token counts (default L).
        Returns text_emb fp32 [n_tags, seq_length, 768]."""
        n, L = tokens.shape
        if L > self.seq_length:
            raise ValueError(f"{L} tokens exceed seq_length {self.seq_length}")
        dev = tokens.device
        src = torch.full((n, self.seq_length), self.pad_id, dtype=torch.int64, device=dev)
        src[:, :L] = tokens
        pos = torch.arange(self.seq_length, device=dev).unsqueeze(0)
        lens = torch.full((n,), L, device=dev) if lengths is None else lengths.to(dev)
        seg = (pos < lens.unsqueeze(1)).to(torch.int64)
        return self.text(src, None, seg)

    @torch.no_grad()
    def image_features(self, frames):
        """frames fp32 [n_img, 3, 224, 224] -> img_emb fp32 [1, n_img, 768] (CLS position of each keyframe)."""
        n = frames.shape[0]
        seg = torch.ones(n, 197, dtype=torch.int64, device=frames.device)
        hidden = self.vit(frames, None, seg)
        return hidden[:, 0, :].unsqueeze(0).contiguous()

    @torch.no_grad()
    def __call__(self, frames, tokens, lengths=None):
        return self.text_features(tokens, lengths), self.image_features(frames)

    def pad_images(self, img_emb, generator=None, shuffle=True):
        """[1, n_img, 768] (or [n_img, 768]) -> [max_imgs, 768]: shuffle, truncate or cyclically repeat
        (finetune/pointwise.py:139-154)."""
        load = img_emb[0] if img_emb.dim() == 3 else img_emb
        n = load.shape[0]
        if shuffle:
            load = load[torch.randperm(n, generator=generator).to(load.device)]       # host permutation, as the dataset
        if n > self.max_imgs:
            return load[:self.max_imgs].clone()
        idx = torch.arange(self.max_imgs, device=load.device) % n
        return load[idx]
