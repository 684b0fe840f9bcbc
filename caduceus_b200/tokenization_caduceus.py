"""Character-level DNA tokenizer with the reference's id assignment and complement map
(ref:caduceus/tokenization_caduceus.py:49-66): specials 0..6, then A7 C8 G9 T10 N11; complement 7<->10, 8<->9.
It is the source of `complement_map` that the trainer injects into the model config (ref:train.py:194-197).
"""
from transformers import PreTrainedTokenizer

_SPECIALS = ("[CLS]", "[SEP]", "[BOS]", "[MASK]", "[PAD]", "[RESERVED]", "[UNK]")


class CaduceusTokenizer(PreTrainedTokenizer):
    model_input_names = ["input_ids"]

    def __init__(self, model_max_length, characters=("A", "C", "G", "T", "N"), complement_map=None,
                 bos_token="[BOS]", eos_token="[SEP]", sep_token="[SEP]", cls_token="[CLS]", pad_token="[PAD]",
                 mask_token="[MASK]", unk_token="[UNK]", **kwargs):
        pairs = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"} if complement_map is None else complement_map
        self.characters = characters
        self.model_max_length = model_max_length
        names = list(_SPECIALS) + list(characters)
        self._vocab_str_to_int = {tok: i for i, tok in enumerate(names)}
        self._vocab_int_to_str = dict(enumerate(names))
        # id -> id of the complementary base; identity for everything without a partner
        self._complement_map = {i: self._vocab_str_to_int[pairs[tok]] if tok in pairs else i
                                for tok, i in self._vocab_str_to_int.items()}
        super().__init__(
            bos_token=bos_token, eos_token=eos_token, sep_token=sep_token, cls_token=cls_token, pad_token=pad_token,
            mask_token=mask_token, unk_token=unk_token, add_prefix_space=kwargs.pop("add_prefix_space", False),
            model_max_length=model_max_length, padding_side=kwargs.pop("padding_side", "left"), **kwargs)

    @property
    def vocab_size(self):
        return len(self._vocab_str_to_int)

    @property
    def complement_map(self):
        return self._complement_map

    def get_vocab(self):
        return self._vocab_str_to_int

    def _tokenize(self, text, **kwargs):
        return list(text.upper())            # soft-masked (lower-case) bases are folded to upper case

    def _convert_token_to_id(self, token):
        return self._vocab_str_to_int.get(token, self._vocab_str_to_int["[UNK]"])

    def _convert_id_to_token(self, index):
        return self._vocab_int_to_str[index]

    def convert_tokens_to_string(self, tokens):
        return "".join(tokens)

    def build_inputs_with_special_tokens(self, token_ids_0, token_ids_1=None):
        # each segment is closed by [SEP]; no [CLS] is prepended
        out = list(token_ids_0) + [self.sep_token_id]
        if token_ids_1 is not None:
            out += list(token_ids_1) + [self.sep_token_id]
        return out

    def get_special_tokens_mask(self, token_ids_0, token_ids_1=None, already_has_special_tokens=False):
        if already_has_special_tokens:
            return super().get_special_tokens_mask(token_ids_0=token_ids_0, token_ids_1=token_ids_1,
                                                   already_has_special_tokens=True)
        mask = [0] * len(token_ids_0) + [1]
        if token_ids_1 is not None:
            mask += [0] * len(token_ids_1) + [1]
        return mask

    def save_vocabulary(self, save_directory, filename_prefix=None):
        return ()                            # fixed vocabulary, nothing to write
