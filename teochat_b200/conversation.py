"""Prompt templates of the hot path.

Behavioural mirror of ``videollava/conversation.py``: ``Conversation.get_prompt`` for the
separator styles the eval path can select (:29-104; TEOChat uses ``v1`` = ``conv_vicuna_v1``
:252-262 with style TWO :51-60) and the ``conv_templates`` registry (:361-377).
"""
from __future__ import annotations

import dataclasses
from enum import Enum, auto
from typing import List, Optional, Sequence


class SeparatorStyle(Enum):
    SINGLE = auto()
    TWO = auto()
    PLAIN = auto()
    LLAMA_2 = auto()


def _text(message):
    # image-carrying messages are (text, image, mode) tuples in the demo (conversation.py:56-57)
    return message[0] if isinstance(message, tuple) else message


@dataclasses.dataclass
class Conversation:
    system: str
    roles: Sequence[str]
    messages: List[List[Optional[str]]]
    offset: int = 0
    sep_style: SeparatorStyle = SeparatorStyle.SINGLE
    sep: str = "###"
    sep2: Optional[str] = None
    version: str = "Unknown"

    def append_message(self, role, message):
        self.messages.append([role, message])

    def copy(self) -> "Conversation":
        return Conversation(system=self.system, roles=self.roles,
                            messages=[[r, m] for r, m in self.messages], offset=self.offset,
                            sep_style=self.sep_style, sep=self.sep, sep2=self.sep2,
                            version=self.version)

    def _messages(self):
        """conversation.py:31-42: when the first message carries an image (a (text, image, mode) tuple, the demo's form),
        its ``<image>`` tag is moved to the front (``<image>\n`` + text), or — for ``mmtag`` versions — sent as a separate
        ``<Image><image></Image>`` / ``Received.`` exchange."""
        messages = self.messages
        if len(messages) > 0 and type(messages[0][1]) is tuple:
            messages = [list(m) for m in self.messages]
            init_role, init_msg = messages[0]
            init_msg = init_msg[0].replace("<image>", "").strip()
            if "mmtag" in self.version:
                messages[0] = [init_role, init_msg]
                messages.insert(0, [self.roles[0], "<Image><image></Image>"])
                messages.insert(1, [self.roles[1], "Received."])
            else:
                messages[0] = [init_role, "<image>\n" + init_msg]
        return messages

    def get_prompt(self) -> str:
        style = self.sep_style
        messages = self._messages()
        if style == SeparatorStyle.SINGLE:
            out = self.system + self.sep
            for role, message in messages:
                out += (role + ": " + _text(message) + self.sep) if message else (role + ":")
            return out
        if style == SeparatorStyle.TWO:
            seps = (self.sep, self.sep2)
            out = self.system + seps[0]
            for i, (role, message) in enumerate(messages):
                out += (role + ": " + _text(message) + seps[i % 2]) if message else (role + ":")
            return out
        if style == SeparatorStyle.PLAIN:
            seps = (self.sep, self.sep2)
            out = self.system
            for i, (_, message) in enumerate(messages):
                if message:
                    out += _text(message) + seps[i % 2]
            return out
        if style == SeparatorStyle.LLAMA_2:
            out = ""
            for i, (role, message) in enumerate(messages):
                if i == 0:
                    assert message, "first message should not be none"
                    assert role == self.roles[0], "first message should come from user"
                if not message:
                    continue
                message = _text(message)
                if i == 0:
                    message = f"<<SYS>>\n{self.system}\n<</SYS>>\n\n" + message
                if i % 2 == 0:
                    out += self.sep + f"[INST] {message} [/INST]"
                else:
                    out += " " + message + " " + self.sep2
            return out.lstrip(self.sep)
        raise ValueError(f"Invalid style: {style}")


conv_vicuna_v1 = Conversation(
    system="A chat between a curious user and an artificial intelligence assistant. "
           "The assistant gives helpful, detailed, and polite answers to the user's questions.",
    roles=("USER", "ASSISTANT"), version="v1", messages=[], offset=0,
    sep_style=SeparatorStyle.TWO, sep=" ", sep2="</s>")

conv_llama_2 = Conversation(
    system="You are a helpful language and vision assistant. You are able to understand the visual "
           "content that the user provides, and assist the user with a variety of tasks using "
           "natural language.",
    roles=("USER", "ASSISTANT"), version="llama_v2", messages=[], offset=0,
    sep_style=SeparatorStyle.LLAMA_2, sep="<s>", sep2="</s>")

conv_llava_plain = Conversation(
    system="", roles=("", ""), messages=[], offset=0, sep_style=SeparatorStyle.PLAIN, sep="\n")

default_conversation = conv_vicuna_v1
conv_templates = {
    "default": conv_vicuna_v1,
    "v1": conv_vicuna_v1,
    "vicuna_v1": conv_vicuna_v1,
    "llava_v1": conv_vicuna_v1,
    "llava_llama_2": conv_llama_2,
    "plain": conv_llava_plain,
}
