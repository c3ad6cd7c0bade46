"""thunder_speech_b200: B200-native (sm_100a) implementation of the thunder-speech ASR forward hot path.

Public surface mirrors the reference's module names for this path only:

    thunder_speech_b200.quartznet.transform.FilterbankFeatures
    thunder_speech_b200.quartznet.blocks.{MaskedConv1d, QuartznetBlock, QuartznetEncoder}
    thunder_speech_b200.citrinet.blocks.{SqueezeExcite, CitrinetBlock, CitrinetEncoder}
    thunder_speech_b200.blocks.{MultiSequential, Masked, lengths_to_mask, get_same_padding, conv1d_decoder}
    thunder_speech_b200.module.CTCModule (= BaseCTCModule forward / predict)
    thunder_speech_b200.text_processing.{Vocabulary, BatchTextTransformer}

All arithmetic runs in ``libthunder_b200.so`` (hand-written CUDA, C ABI in include/thunder_b200.h) through the
``torch.ops.thunder_b200.*`` custom ops; there is no CPU or eager fallback.
"""
__version__ = "0.1.0"
